"""Screening: the Polyak iteration on the induced vector potential
inside every time step (reference solver/solver.py:522-578, 650-688; all-pairs kernel
solver/screening.py:12-42; site average finite_volume/mesh.py:203-243) on the device, against
the oracle's restatement, which tests/test_oracle_vs_reference.py pins to the reference's own
numba kernel.  Tolerance: 1e-8 gauge-fixed on psi, mu, J_s, J_n; 1e-8 on A_induced; the number
of Polyak passes of every step and the dt sequence must be equal."""
import numpy as np
import pytest

from oracle import tdgl_oracle as orc

pytestmark = pytest.mark.gpu

SCALE = 0.2
OKW = dict(solve_time=0.3, dt_init=1e-4, dt_max=1e-2, include_screening=True,
           screening_tolerance=1e-3, max_iterations_per_step=1000)


@pytest.fixture(scope="module")
def problem():
    from tdgl_b200.synthetic import film_problem

    mesh, A, eps, _ = film_problem(24, 12, 0.4, b=0.4)
    o = orc.OracleSolver(mesh, orc.OracleOptions(**OKW), A, eps, screening_scale=SCALE,
                         probe_points=[10, len(mesh.sites) // 2])
    ref = orc.run(o, end_time=OKW["solve_time"])
    return mesh, A, eps, ref


def _solver(mesh, A, eps, use_graph=True, **over):
    from tdgl_b200 import SolverOptions, TDGLSolver

    kw = dict(OKW, save_every=16, use_cuda_graph=use_graph)
    kw.update(over)
    return TDGLSolver.from_dimensionless(
        mesh, SolverOptions(**kw), A_applied=A, epsilon=eps, screening_scale=SCALE,
        probe_point_indices=[10, len(mesh.sites) // 2])


@pytest.mark.parametrize("use_graph", [True, False])
def test_screening_matches_oracle(problem, use_graph):
    mesh, A, eps, ref = problem
    assert len(mesh.sites) > 1500
    sol = _solver(mesh, A, eps, use_graph).solve()
    d = sol.tdgl_data
    got = dict(psi=d.psi, mu=d.mu, supercurrent=d.supercurrent, normal_current=d.normal_current,
               dt=sol.dynamics.dt)
    assert len(got["dt"]) == ref["steps"]
    np.testing.assert_array_equal(sol.dynamics.screening_iterations, ref["screening_iterations"])
    diff = orc.compare(got, ref, mesh.areas)
    a_ref = ref["induced_vector_potential"]
    diff["A_induced"] = float(np.abs(d.induced_vector_potential - a_ref).max() / np.abs(a_ref).max())
    print("screening", "graph" if use_graph else "host-driven", diff, "passes",
          int(ref["screening_iterations"].sum()), "max", int(ref["screening_iterations"].max()),
          sol.solver_stats)
    for k, v in diff.items():
        assert v < 1e-8, (k, diff)
    np.testing.assert_allclose(got["dt"], ref["dt"], rtol=1e-10)
    assert np.abs(a_ref).max() > 1e-3 and ref["screening_iterations"].max() > 5
    assert sol.solver_stats["screening_iterations"] == int(ref["screening_iterations"].sum())
    if use_graph:
        assert sol.solver_stats["graph_mode"] == 1


def test_screening_failure_raises_like_reference(problem):
    """solver.py:657-663: RuntimeError after max_iterations_per_step passes."""
    mesh, A, eps, _ = problem
    s = _solver(mesh, A, eps, max_iterations_per_step=3)
    with pytest.raises(RuntimeError, match=r"Screening calculation failed to converge at step 0"
                                           r" after 3 iterations\. Relative error in induced"
                                           r" vector potential: .* \(tolerance: 1\.00e-03\)\."):
        s.solve()
    o = orc.OracleSolver(mesh, orc.OracleOptions(**dict(OKW, max_iterations_per_step=3)), A, eps,
                         screening_scale=SCALE)
    with pytest.raises(RuntimeError, match="Screening calculation failed to converge at step 0"):
        orc.run(o, end_time=0.3)


def test_screening_through_the_step_seam(problem):
    """update() threads induced_vector_potential like Runner does (runner.py:417-428)."""
    mesh, A, eps, ref = problem
    s = _solver(mesh, A, eps)
    E = len(mesh.edge_mesh.edges)
    names = ["psi", "mu", "supercurrent", "normal_current", "induced_vector_potential"]
    values = [s.psi_init, s.mu_init, np.zeros(E), np.zeros(E), np.zeros((E, 2))]
    time, dt = 0.0, OKW["dt_init"]

    class Running:
        def __init__(self):
            self.values = {}

        def append(self, k, v):
            self.values.setdefault(k, []).append(v)

    running = Running()
    n = 12
    for i in range(n):
        res = s.update({"step": i, "time": time, "dt": dt}, running, dt, **dict(zip(names, values)))
        new_dt, *values = res
        dt = new_dt
        time += dt
    o = orc.OracleSolver(mesh, orc.OracleOptions(**OKW), A, eps, screening_scale=SCALE)
    r = orc.run(o, end_time=1e9, max_steps=n)
    got = dict(zip(names, values))
    diff = orc.compare(got, r, mesh.areas)
    print("screening via update()", diff)
    for k, v in diff.items():
        assert v < 1e-8, (k, diff)
    np.testing.assert_array_equal(running.values["screening_iterations"], r["screening_iterations"])
    np.testing.assert_allclose(got["induced_vector_potential"], r["induced_vector_potential"],
                               rtol=0, atol=1e-8 * np.abs(r["induced_vector_potential"]).max())
