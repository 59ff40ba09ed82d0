"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (pyTDGL at
/root/reference) through ``oracle/ref_loader.py``.  Run in the build container:

    python oracle/make_golden.py

The fixtures hold the mesh (as produced by the reference's own
``Mesh.from_triangulation``), the dimensionless inputs, per-operator known answers and
whole trajectories, so that the GPU box (which has no /root/reference) can check both the
oracle restatement and the CUDA path against the reference's numbers.
"""

from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader as rl  # noqa: E402
from tdgl_b200.mesh import make_film_points, triangulate  # noqa: E402
from tdgl_b200.synthetic import (  # noqa: E402
    box_terminal, gaussian_disorder, uniform_field_vector_potential)

OUT = os.path.join(ROOT, "tests", "golden")


def mesh_arrays(mesh):
    em = mesh.edge_mesh
    return dict(
        sites=mesh.sites, elements=mesh.elements, boundary_indices=mesh.boundary_indices,
        areas=mesh.areas, edges=em.edges, centers=em.centers, directions=em.directions,
        edge_lengths=em.edge_lengths, dual_edge_lengths=em.dual_edge_lengths,
        boundary_edge_indices=em.boundary_edge_indices)


def operator_vectors(ref, solver, seed):
    """Known answers for each operator of the step on a random state."""
    rng = np.random.default_rng(seed)
    n = len(solver.psi_init)
    ops = solver.operators
    psi = (0.3 + 0.7 * rng.random(n)) * np.exp(2j * np.pi * rng.random(n))
    mu = rng.normal(size=n)
    dt = 1.5e-3
    out = dict(op_psi=psi, op_mu=mu, op_dt=dt)
    out["op_lap_psi"] = ops.psi_laplacian @ psi
    res = ref.TDGLSolver.solve_for_psi_squared(
        psi=psi, abs_sq_psi=np.abs(psi) ** 2, mu=mu, epsilon=solver.epsilon,
        gamma=solver.gamma, u=solver.u, dt=dt, psi_laplacian=ops.psi_laplacian)
    assert res is not None
    out["op_psi_new"], out["op_sq_new"] = res
    js = ops.get_supercurrent(psi)
    out["op_supercurrent"] = js
    mub = rng.normal(size=len(solver.mu_boundary))
    out["op_mu_boundary"] = mub
    out["op_rhs"] = ops.divergence @ js - ops.mu_boundary_laplacian @ mub
    out["op_lap_mu"] = ops.mu_laplacian @ mu
    out["op_normal_current"] = -(ops.mu_gradient @ mu)
    return out


def ramp_field(centers, b_max, t_ramp):
    """Uniform field ramped linearly from 0 to ``b_max`` over ``t_ramp`` (the reference's
    ``LinearRamp * ConstantField``), as a function t -> A[E, 2]."""
    A1 = uniform_field_vector_potential(centers, 1.0)
    return lambda t: min(max(t / t_ramp, 0.0), 1.0) * b_max * A1


def case(name, *, width, height, h, b, disorder, terminals, currents, opts, end_time,
         max_steps, holes=(), probe_xy=None, seed=0, record_every=250, ramp=None):
    ref = rl.load()
    pts = make_film_points(width, height, h, holes=holes, seed=seed)
    pts, tri = triangulate(pts, holes)
    mesh = rl.make_reference_mesh(pts, tri)          # the reference's own mesh arrays
    A = uniform_field_vector_potential(mesh.edge_mesh.centers, b)
    A_func = None
    if ramp is not None:      # (b_max, t_ramp): time-dependent vector potential
        A_func = ramp_field(mesh.edge_mesh.centers, *ramp)
        A = A_func(0.0)
    eps = gaussian_disorder(mesh.sites) if disorder else np.ones(len(mesh.sites))
    terms = ()
    if terminals:
        x0, x1 = -width / 2, width / 2
        tol, big = 1e-9 * width, 10 * max(width, height)
        terms = tuple(sorted(
            (box_terminal(mesh, "source", x0 - tol, x0 + tol, -big, big),
             box_terminal(mesh, "drain", x1 - tol, x1 + tol, -big, big)),
            key=lambda t: t.length))
    probes = None
    if probe_xy is not None:
        probes = [mesh.closest_site(xy) for xy in probe_xy]
    ro = ref.SolverOptions(**opts)
    solver = rl.make_reference_solver(
        mesh, ro, A_applied=A, epsilon=eps,
        terminal_info=[ref.TerminalInfo(*t) for t in terms],
        terminal_currents=currents, probe_points=probes, A_func=A_func)
    data = mesh_arrays(mesh)
    if ramp is not None:
        data["ramp"] = np.asarray(ramp, float)
    data.update(A_applied=A, epsilon=eps, u=solver.u, gamma=solver.gamma)
    data.update(operator_vectors(ref, solver, seed + 1))
    for k, t in enumerate(terms):
        data[f"term{k}_name"] = np.array(t.name)
        data[f"term{k}_sites"] = np.asarray(t.site_indices)
        data[f"term{k}_edges"] = np.asarray(t.edge_indices)
        data[f"term{k}_bedges"] = np.asarray(t.boundary_edge_indices)
        data[f"term{k}_length"] = t.length
        data[f"term{k}_current"] = (currents or {}).get(t.name, 0.0)
    data["n_terminals"] = len(terms)
    if probes is not None:
        data["probe_points"] = np.asarray(probes)
    for k, v in opts.items():
        data[f"opt_{k}"] = v
    data["end_time"] = end_time
    data["max_steps"] = -1 if max_steps is None else max_steps
    res = rl.run_reference(solver, end_time=end_time, max_steps=max_steps,
                           record_every=record_every)
    data.update(psi=res["psi"], mu=res["mu"], supercurrent=res["supercurrent"],
                normal_current=res["normal_current"], dt=res["dt"], steps=res["steps"],
                time=res["time"])
    if probes is not None:
        data["running_mu"] = res["running"]["mu"]
        data["running_theta"] = res["running"]["theta"]
    snaps = res["snapshots"]
    data["snap_steps"] = np.array([s["step"] for s in snaps])
    data["snap_psi"] = np.array([s["psi"] for s in snaps])
    data["snap_mu"] = np.array([s["mu"] for s in snaps])
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **data)
    print(f"{name}: N={len(mesh.sites)} E={len(mesh.edge_mesh.edges)} steps={res['steps']}"
          f" time={res['time']:.4f} |psi| in [{np.abs(res['psi']).min():.3f},"
          f" {np.abs(res['psi']).max():.3f}] -> {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    only = set(sys.argv[1:])
    _case = case

    def case(name, **kw):  # noqa: F811  (optional: regenerate only the named cases)
        if not only or name in only:
            _case(name, **kw)

    # config-1 geometry with a deterministic perturbation so every term is exercised
    # (SURVEY.md §8c): fixed dt, 1000 steps
    case("film20_fixed", width=20, height=20, h=0.35, b=0.3, disorder=True,
         terminals=False, currents=None,
         opts=dict(solve_time=1e9, adaptive=False, dt_init=2e-3, dt_max=2e-3),
         end_time=1e9, max_steps=1000, probe_xy=[(-5.0, 0.0), (5.0, 0.0)])
    # adaptive time step, vortex entry in a field
    case("film20_adaptive", width=20, height=20, h=0.35, b=0.4, disorder=True,
         terminals=False, currents=None,
         opts=dict(solve_time=20.0, dt_init=1e-4, dt_max=1e-1),
         end_time=20.0, max_steps=None, probe_xy=[(-5.0, 0.0), (5.0, 0.0)])
    # time-dependent vector potential: field ramped from 0 to 0.3 Bc2 over 2 time units,
    # fixed dt (smooth regime: parity to roundoff-level tolerances at the end)
    case("film20_ramp", width=20, height=20, h=0.35, b=0.0, disorder=True,
         terminals=False, currents=None,
         opts=dict(solve_time=1e9, adaptive=False, dt_init=2e-3, dt_max=2e-3),
         end_time=1e9, max_steps=600, probe_xy=[(-5.0, 0.0), (5.0, 0.0)], ramp=(0.3, 2.0))
    # transport: terminals + holes + current, adaptive
    case("strip_transport", width=40, height=10, h=0.5, b=0.05, disorder=False,
         terminals=True, currents={"source": 2.0, "drain": -2.0},
         holes=((-8.0, 0.0, 2.0), (9.0, 1.0, 1.5)),
         opts=dict(solve_time=10.0, dt_init=1e-4, dt_max=1e-1),
         end_time=10.0, max_steps=None, probe_xy=[(-15.0, 0.0), (15.0, 0.0)],
         record_every=50)  # a symmetry-breaking instability amplifies roundoff x100 per 10
    #                        steps from step ~170 on: parity is asserted at step 150
