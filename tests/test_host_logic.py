"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host-side
setup (C++ AMG hierarchy) is sound, and the Python mirror of the reference interface
behaves like the reference (option validation, errors, Runner bookkeeping containers)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

import tdgl_b200 as tdgl
from tdgl_b200 import _lib
from tdgl_b200.engine import host_amg_probe
from tdgl_b200.mesh import make_film_mesh
from tdgl_b200.solution import SavedSteps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "tdgl_b200.h")).read()
    declared = set(re.findall(r"\b(tdgl_[A-Za-z_0-9]+)\s*\(", header))
    declared -= {"tdgl_handle", "tdgl_config", "tdgl_advance_info"}
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.tdgl_version()
    assert ctypes.sizeof(_lib.tdgl_config) == 64
    assert ctypes.sizeof(_lib.tdgl_advance_info) == 112


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    mesh = make_film_mesh(6, 6, 0.5)
    with pytest.raises(_lib.TDGLLibraryError):
        tdgl.DeviceEngine(mesh)


def test_host_amg_hierarchy_converges():
    mesh = make_film_mesh(60, 40, 0.43, holes=((5.0, 3.0, 6.0),))
    n = len(mesh.sites)
    rng = np.random.default_rng(0)
    b = rng.normal(size=n)
    b -= b.mean()
    r = host_amg_probe(mesh, rhs=b, rtol=1e-10)
    assert r["levels"] >= 3 and r["rows"][0] == n and r["rows"][-1] <= 200
    assert sum(r["nnz"]) / r["nnz"][0] < 1.6          # operator complexity
    assert 0 < r["iterations"] <= 30, r["iterations"]
    # residual of the symmetrised system
    em = mesh.edge_mesh
    w = em.dual_edge_lengths / em.edge_lengths
    x = r["x"]
    Ax = np.zeros(n)
    d = x[em.edges[:, 0]] - x[em.edges[:, 1]]
    np.add.at(Ax, em.edges[:, 0], w * d)
    np.add.at(Ax, em.edges[:, 1], -w * d)
    assert np.linalg.norm(b - Ax) / np.linalg.norm(b) < 1e-9


def test_solver_options_validation_matches_reference():
    with pytest.raises(tdgl.SolverOptionsError, match="dt_init must be less than"):
        tdgl.SolverOptions(solve_time=1, dt_init=1.0, dt_max=0.1).validate()
    with pytest.raises(tdgl.SolverOptionsError, match="terminal_psi"):
        tdgl.SolverOptions(solve_time=1, terminal_psi=2.0).validate()
    with pytest.raises(tdgl.SolverOptionsError, match="adaptive_time_step_multiplier"):
        tdgl.SolverOptions(solve_time=1, adaptive_time_step_multiplier=1.5).validate()
    with pytest.raises(tdgl.SolverOptionsError, match="sparse solver must be one of"):
        tdgl.SolverOptions(solve_time=1, sparse_solver="magic").validate()
    o = tdgl.SolverOptions(solve_time=1, sparse_solver="superlu")
    o.validate()
    assert o.sparse_solver is tdgl.SparseSolver.SUPERLU
    assert issubclass(tdgl.SolverOptionsError, ValueError)
    # the reference's 24 fields come first, in order, with the same defaults
    import dataclasses

    names = [f.name for f in dataclasses.fields(tdgl.SolverOptions)][:24]
    assert names == [
        "solve_time", "skip_time", "dt_init", "dt_max", "adaptive", "adaptive_window",
        "max_solve_retries", "adaptive_time_step_multiplier", "output_file", "terminal_psi",
        "gpu", "sparse_solver", "pause_on_interrupt", "save_every", "progress_interval",
        "monitor", "monitor_update_interval", "field_units", "current_units",
        "include_screening", "max_iterations_per_step", "screening_tolerance",
        "screening_step_size", "screening_step_drag"]


def test_device_terminals_and_units():
    layer = tdgl.Layer(london_lambda=2.0, coherence_length=0.5, thickness=0.1)
    film = tdgl.Polygon("film", points=tdgl.box(10, 4))
    # (thin terminal polygons: boundary edges belong to a terminal when their CENTRES lie in
    # it, ref device.py:243-245 — the horizontal edges next to the corners must stay out)
    src = tdgl.Polygon("source", points=tdgl.box(0.02, 4, center=(-5, 0)))
    drn = tdgl.Polygon("drain", points=tdgl.box(0.02, 4, center=(5, 0)))
    hole = tdgl.Polygon("hole", points=tdgl.circle(0.8, points=40))
    dev = tdgl.Device("bar", layer=layer, film=film, holes=[hole], terminals=[src, drn],
                      probe_points=[(-3, 0), (3, 0)])
    dev.make_mesh(max_edge_length=0.25)
    mesh = dev.mesh
    assert len(mesh.sites) > 500
    assert abs(mesh.areas.sum() * 0.25 - (40 - np.pi * 0.8**2)) < 0.05   # areas in xi^2
    info = dev.terminal_info()
    assert [t.name for t in info] in (["source", "drain"], ["drain", "source"])
    for t in info:
        assert abs(t.length - 4.0) < 1e-9
        assert len(t.site_indices) == len(t.boundary_edge_indices) + 1
    # Bc2 = Phi0 / (2 pi xi^2) with xi = 0.5 um
    assert abs(dev.Bc2 - 2.067833848e-15 / (2 * np.pi * (0.5e-6) ** 2)) < 1e-12
    assert len(dev.probe_point_indices) == 2
    assert dev == dev.copy()


def test_make_mesh_contract_of_the_reference():
    """Device.make_mesh (ref device.py:520-566, meshing.py:15-123): the result has at least
    `min_points` vertices and no edge longer than `max_edge_length` (default 1.0 x xi);
    `smooth` runs Laplacian sweeps that keep the boundary; same seed, same mesh."""
    layer = tdgl.Layer(london_lambda=2.0, coherence_length=0.5, thickness=0.1)
    film = tdgl.Polygon("film", points=tdgl.box(10, 4))
    hole = tdgl.Polygon("hole", points=tdgl.circle(0.8, points=40))

    def device():
        return tdgl.Device("bar", layer=layer, film=film, holes=[hole])

    def longest(mesh):      # in length units (mesh sites are in units of xi)
        return mesh.edge_mesh.edge_lengths.max() * layer.coherence_length

    d = device()
    d.make_mesh()                                         # default: max_edge_length = xi
    assert longest(d.mesh) <= 0.5 + 1e-12
    n_default = len(d.mesh.sites)
    d.make_mesh(max_edge_length=0.2)
    assert longest(d.mesh) <= 0.2 + 1e-12 and len(d.mesh.sites) > 2 * n_default
    d.make_mesh(max_edge_length=1.0, min_points=3000)     # min_points decides
    assert len(d.mesh.sites) >= 3000 and longest(d.mesh) <= 1.0
    d.make_mesh(max_edge_length=0.1, min_points=100)      # max_edge_length decides
    assert longest(d.mesh) <= 0.1 + 1e-12
    d.make_mesh(max_edge_length=-1)                       # polygon point density only
    seg = 2 * np.pi * 0.8 / 40
    assert 0.3 * seg < np.median(d.mesh.edge_mesh.edge_lengths) * 0.5 < 3 * seg
    # smoothing: boundary sites stay, interior sites move, areas still tile the film
    a, b = device(), device()
    a.make_mesh(max_edge_length=0.25, reorder=False)
    b.make_mesh(max_edge_length=0.25, smooth=20, reorder=False)
    assert a.mesh.elements.shape == b.mesh.elements.shape
    bi = a.mesh.boundary_indices
    assert np.array_equal(a.mesh.sites[bi], b.mesh.sites[bi])
    assert np.abs(a.mesh.sites - b.mesh.sites).max() > 1e-3
    assert abs(b.mesh.areas.sum() * 0.25 - (40 - np.pi * 0.8**2)) < 0.05
    # (a smoothed jittered lattice is more uniform: smaller spread of the edge lengths)
    assert b.mesh.edge_mesh.edge_lengths.std() < a.mesh.edge_mesh.edge_lengths.std()
    c = device()
    c.make_mesh(max_edge_length=0.25, smooth=20, reorder=False)
    assert np.array_equal(b.mesh.sites, c.mesh.sites)


def test_saved_steps_dynamics_matches_reference_layout():
    s = SavedSteps()
    s.save_fixed_values({"epsilon": np.ones(4)})
    s.save_time_step({"step": 0, "time": 0.0, "dt": 1e-3}, {"psi": np.ones(4, complex)}, None)
    rs = {"dt": np.array([[1e-3, 2e-3, 0.0]]), "mu": np.array([[1.0, 2.0, 0.0], [3.0, 4.0, 0.0]])}
    s.save_time_step({"step": 2, "time": 3e-3, "dt": 2e-3}, {"psi": np.ones(4, complex)}, rs)
    dyn = s.dynamics()
    np.testing.assert_allclose(dyn.dt, [1e-3, 2e-3])
    np.testing.assert_allclose(dyn.time, [1e-3, 3e-3])
    np.testing.assert_allclose(dyn.voltage(0, 1), [-2.0, -2.0])


def _small_problem():
    from tdgl_b200.synthetic import film_problem

    return film_problem(8, 4, 0.5, b=0.1, terminals=True)


def test_solver_input_errors_match_reference():
    """The errors ``TDGLSolver.__init__`` raises before any device work, with the reference's
    messages (solver.py:35-60, 186-189, 215-216, 228-232; tdgl/test/test_solve.py:34-79)."""
    mesh, A, eps, terms = _small_problem()
    opts = tdgl.SolverOptions(solve_time=1.0)
    mk = tdgl.TDGLSolver.from_dimensionless
    with pytest.raises(ValueError, match="epsilon must be <= 1"):
        mk(mesh, opts, A_applied=A, epsilon=2.0 * eps, terminal_info=terms)
    with pytest.raises(ValueError, match=r"Unknown terminal\(s\)"):
        mk(mesh, opts, A_applied=A, epsilon=eps, terminal_info=terms,
           terminal_currents={"source": 1.0, "nowhere": -1.0})
    with pytest.raises(ValueError, match="sum of all terminal currents must be 0"):
        mk(mesh, opts, A_applied=A, epsilon=eps, terminal_info=terms,
           terminal_currents={"source": 1.0, "drain": -0.5})
    with pytest.raises(ValueError, match="sum of all terminal currents must be 0"):
        mk(mesh, opts, A_applied=A, epsilon=eps, terminal_info=terms,
           terminal_currents=lambda t: {"source": 1.0 + t, "drain": -1.0})
    empty = terms[0]._replace(name="ghost", length=0.0)
    with pytest.raises(ValueError, match="does not contain any points"):
        mk(mesh, opts, A_applied=A, epsilon=eps, terminal_info=(empty,))
    with pytest.raises(tdgl.SolverOptionsError):
        mk(mesh, tdgl.SolverOptions(solve_time=1.0, dt_init=1.0, dt_max=0.1), A_applied=A,
           epsilon=eps)
    # screening is an all-pairs sum over the whole mesh: rejected loudly on a sharded job
    with pytest.raises(tdgl.SolverOptionsError, match="include_screening"):
        mk(mesh, tdgl.SolverOptions(solve_time=1.0, include_screening=True, distributed=True),
           A_applied=A, epsilon=eps)
    for bad in (dict(screening_step_drag=0.0), dict(screening_step_size=0.0),
                dict(screening_tolerance=0.0)):                  # options.py:120-135
        with pytest.raises(tdgl.SolverOptionsError):
            mk(mesh, tdgl.SolverOptions(solve_time=1.0, include_screening=True, **bad),
               A_applied=A, epsilon=eps)


def test_mesh_edge_cases():
    """Mesh input contract (reference finite_volume/mesh.py:104-151): shape errors, boundary
    edges = edges of exactly one triangle, areas tile the film."""
    from tdgl_b200.mesh import Mesh

    with pytest.raises(ValueError, match=r"shape \(n, 2\)"):
        Mesh.from_triangulation(np.zeros((4, 3)), np.array([[0, 1, 2]]))
    with pytest.raises(ValueError, match=r"shape \(m, 3\)"):
        Mesh.from_triangulation(np.zeros((4, 2)), np.array([[0, 1, 2, 3]]))
    # the smallest mesh: two triangles of the unit square
    m = Mesh.from_triangulation(np.array([[0, 0], [1, 0], [1, 1], [0, 1.0]]),
                                np.array([[0, 1, 2], [0, 2, 3]]))
    em = m.edge_mesh
    assert len(em.edges) == 5 and len(em.boundary_edge_indices) == 4
    assert sorted(m.boundary_indices) == [0, 1, 2, 3]
    assert np.all(em.edges[:, 0] < em.edges[:, 1])
    # values of the reference's own Mesh.from_triangulation on this mesh (its Voronoi cells
    # of corner sites whose circumcentres fall on the hypotenuse do not tile the square)
    np.testing.assert_allclose(m.areas, [0.125, 0.25, 0.125, 0.25])
    np.testing.assert_allclose(em.dual_edge_lengths, [0.5, 0.0, 0.5, 0.5, 0.5])
    mesh = make_film_mesh(12, 7, 0.5, holes=((1.0, 0.5, 2.0),))
    assert abs(mesh.areas.sum() - (12 * 7 - np.pi * 2.0**2)) < 0.25   # polygonal hole outline
    assert np.all(mesh.areas > 0) and np.all(mesh.edge_mesh.dual_edge_lengths >= 0)


def test_sources_ramp_times_field_is_separable():
    """``LinearRamp * ConstantField`` (reference sources/scaling.py, sources/constant.py)."""
    f = tdgl.LinearRamp(tmin=1.0, tmax=3.0, initial=0.5, final=2.0) * tdgl.ConstantField(0.2)
    x, y, z = np.array([0.0, 1.0, 2.0]), np.array([0.0, 0.0, 1.0]), np.zeros(3)
    A0 = tdgl.ConstantField(0.2)(x, y, z)
    assert f.time_dependent and f.separable is not None
    for t, s in [(0.0, 0.5), (1.0, 0.5), (2.0, 1.25), (3.0, 2.0), (9.0, 2.0)]:
        np.testing.assert_allclose(f(x, y, z, t=t), s * A0, rtol=0, atol=1e-15)
    assert not tdgl.ConstantField(0.2).time_dependent
    with pytest.raises(ValueError):
        tdgl.LinearRamp(tmin=1.0, tmax=1.0)


def test_solution_npz_round_trip(tmp_path):
    """Save / load equality of the output tree (reference test_solution.py:53-61, with the
    .npz stand-in for HDF5: h5py is not in this image)."""
    from tdgl_b200.solution import Solution

    n, e = 5, 7
    rng = np.random.default_rng(0)
    s = SavedSteps()
    s.save_fixed_values({"applied_vector_potential": rng.normal(size=(e, 2)), "epsilon": np.ones(n)})
    for k in range(3):
        vals = {"psi": rng.normal(size=n) + 1j * rng.normal(size=n), "mu": rng.normal(size=n),
                "supercurrent": rng.normal(size=e), "normal_current": rng.normal(size=e),
                "induced_vector_potential": np.zeros((e, 2))}
        rs = None if k == 0 else {"dt": np.array([[1e-3, 2e-3]]),
                                  "mu": rng.normal(size=(2, 2)), "theta": rng.normal(size=(2, 2))}
        s.save_time_step({"step": 2 * k, "time": 3e-3 * k, "dt": 2e-3}, vals, rs)
    opts = tdgl.SolverOptions(solve_time=1.0)
    sol = Solution(device=None, options=opts, saved=s)
    sol.load_tdgl_data(-1)
    path = sol.to_npz(str(tmp_path / "out"))
    back = Solution.from_npz(path, options=opts)
    assert back.data_range == sol.data_range == (0, 2)
    for step in range(3):
        sol.solve_step = back.solve_step = step
        a, b = sol.tdgl_data, back.tdgl_data
        for name in ("psi", "mu", "supercurrent", "normal_current", "applied_vector_potential",
                     "induced_vector_potential", "epsilon"):
            np.testing.assert_array_equal(getattr(a, name), getattr(b, name))
        assert a.state["step"] == b.state["step"] and a.state["time"] == b.state["time"]
    np.testing.assert_array_equal(sol.dynamics.dt, back.dynamics.dt)
    np.testing.assert_array_equal(sol.dynamics.mu, back.dynamics.mu)
    np.testing.assert_array_equal(sol.dynamics.voltage(0, 1), back.dynamics.voltage(0, 1))


def test_fp32_vcycle_keeps_iteration_count():
    """Design study for the next round, on the SciPy prototype of the mu
    solver (oracle/amg_proto.py): running the V-cycle — the preconditioner only — in fp32
    leaves the fp64 CG at the same iteration count and the same final residual."""
    from oracle import amg_proto as ap

    mesh = make_film_mesh(80, 50, 0.43, holes=((5.0, 3.0, 6.0),))
    n = len(mesh.sites)
    em = mesh.edge_mesh
    A = ap.sym_mu_matrix(em.edges, em.edge_lengths, em.dual_edge_lengths, n).tocsr()
    levels = ap.build_hierarchy(A, theta=0.08)
    b = A @ np.random.default_rng(0).normal(size=n)
    _, h64 = ap.pcg(A, b, lambda r: ap.vcycle(levels, r))
    lv32 = []
    for lv in levels:
        m = ap.Level()
        m.A, m.d, m.rho = lv.A.astype(np.float32), lv.d.astype(np.float32), lv.rho
        if hasattr(lv, "P"):
            m.P, m.R = lv.P.astype(np.float32), lv.R.astype(np.float32)
        if hasattr(lv, "pinv"):
            m.pinv = lv.pinv.astype(np.float32)
        lv32.append(m)

    def m32(r):
        s = np.abs(r).max()
        return ap.vcycle(lv32, (r / s).astype(np.float32)).astype(np.float64) * s

    _, h32 = ap.pcg(A, b, m32)
    assert h64[-1] < 1e-10 and h32[-1] < 1e-10
    assert abs(len(h32) - len(h64)) <= 1, (len(h32), len(h64))


def test_c_abi_from_plain_c(tmp_path):
    """include/tdgl_b200.h is a plain C header and the library links from a C program
    (the boundary a non-Python host would bind): version string, struct sizes and the
    argument check of tdgl_create, which happens before any device work."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    _lib.load()
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include "tdgl_b200.h"
#include <stdio.h>
#include <string.h>
int main(void) {
  tdgl_config cfg; memset(&cfg, 0, sizeof cfg); cfg.struct_size = (int32_t)sizeof cfg;
  printf("%s %zu %zu\n", tdgl_version(), sizeof cfg, sizeof(tdgl_advance_info));
  tdgl_handle* h = 0;
  int rc = tdgl_create(&h, 2, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 1.0, 0, 0, &cfg);
  printf("%d %s\n", rc, tdgl_last_error(0));
  return rc == TDGL_E_INVALID && h == 0 ? 0 : 1;
}
''')
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-ltdgl_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out and " 64 112" in out and "null array argument" in out


def test_solution_npz_round_trip_keeps_options_and_mesh(tmp_path):
    """A saved Solution can be loaded without arguments: options (save_every feeds
    ``Solution.times``) and the mesh come back from the file; the tree of keys is the
    reference's (data/<k>/..., running_state, mesh/, solution/options)."""
    from tdgl_b200.solution import Solution

    mesh = make_film_mesh(6, 4, 0.5)
    n, E = len(mesh.sites), len(mesh.edge_mesh.edges)
    saved = SavedSteps()
    saved.save_fixed_values({"epsilon": np.ones(n), "applied_vector_potential": np.zeros((E, 2))})
    rng = np.random.default_rng(0)
    for k in range(3):
        vals = {"psi": rng.normal(size=n) + 1j * rng.normal(size=n), "mu": rng.normal(size=n),
                "supercurrent": rng.normal(size=E), "normal_current": rng.normal(size=E),
                "induced_vector_potential": np.zeros((E, 2))}
        rs = None if k == 0 else {"dt": np.full((1, 2), 1e-3), "mu": np.ones((2, 2))}
        saved.save_time_step({"step": 2 * k, "time": 2e-3 * k, "dt": 1e-3}, vals, rs)
    opts = tdgl.SolverOptions(solve_time=1.0, save_every=2, terminal_psi=None)
    sol = Solution(device=None, options=opts, saved=saved, mesh=mesh, total_seconds=1.5)
    path = sol.to_npz(str(tmp_path / "out.npz"))
    with np.load(path) as f:
        keys = set(f.files)
    assert {"epsilon", "data/0/psi", "data/2/running_state/dt", "data/1/attrs/step",
            "mesh/sites", "mesh/edge_mesh/edges", "solution/options"} <= keys
    back = Solution.from_npz(path)
    assert back.options.save_every == 2 and back.options.terminal_psi is None
    np.testing.assert_array_equal(back.tdgl_data.psi, sol.tdgl_data.psi)
    np.testing.assert_array_equal(back._mesh.edge_mesh.edges, mesh.edge_mesh.edges)
    np.testing.assert_allclose(back.times, sol.times)
    assert back.total_seconds == 1.5
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            sol.to_hdf5(str(tmp_path / "out.h5"))


class _FakeH5Group(dict):
    """The slice of the h5py API Solution.to_hdf5 uses (h5py / libhdf5 are not in this image):
    groups are dicts with an ``attrs`` dict, ``f["a/b/c"] = array`` creates the parents."""

    def __init__(self):
        super().__init__()
        self.attrs = {}

    def require_group(self, path):
        g = self
        for part in path.split("/"):
            if part not in g:
                dict.__setitem__(g, part, _FakeH5Group())
            g = dict.__getitem__(g, part)
        return g

    def __setitem__(self, path, value):
        *parents, leaf = path.split("/")
        g = self.require_group("/".join(parents)) if parents else self
        dict.__setitem__(g, leaf, np.asarray(value))

    def __getitem__(self, path):
        g = self
        for part in path.split("/"):
            g = dict.__getitem__(g, part)
        return g


def test_solution_to_hdf5_writes_the_reference_layout(tmp_path, monkeypatch):
    """Solution.to_hdf5 against a stand-in for h5py: the layout of the reference's files —
    `data/<k>` groups carrying step / time / dt as ATTRIBUTES and the fields as datasets,
    `data/<k>/running_state/<name>` (runner.py:155-183), `mesh/...` (mesh.py:345-368), solver
    options as attributes of `solution/options` (solution.py:874-931)."""
    import importlib.machinery
    import types

    from tdgl_b200.solution import Solution

    files = {}

    class File(_FakeH5Group):
        def __init__(self, path, mode):
            super().__init__()
            assert mode == "x"
            files[path] = self

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    fake = types.ModuleType("h5py")
    fake.File = File
    fake.__spec__ = importlib.machinery.ModuleSpec("h5py", None)
    monkeypatch.setitem(sys.modules, "h5py", fake)

    mesh = make_film_mesh(6, 4, 0.5)
    n, E = len(mesh.sites), len(mesh.edge_mesh.edges)
    saved = SavedSteps()
    saved.save_fixed_values({"epsilon": np.ones(n)})
    for k in range(2):
        vals = {"psi": np.full(n, 1.0 + 1j * k), "mu": np.full(n, float(k)),
                "supercurrent": np.zeros(E), "normal_current": np.zeros(E)}
        rs = None if k == 0 else {"dt": np.full((1, 3), 1e-3), "mu": np.ones((2, 3))}
        saved.save_time_step({"step": 3 * k, "time": 3e-3 * k, "dt": 1e-3}, vals, rs)
    opts = tdgl.SolverOptions(solve_time=1.0, save_every=3)
    sol = Solution(device=None, options=opts, saved=saved, mesh=mesh, total_seconds=2.0)
    path = sol.to_hdf5(str(tmp_path / "out.h5"))
    f = files[path]
    assert set(f) >= {"data", "mesh", "solution", "epsilon"}
    assert set(f["data"]) == {"0", "1"}
    g1 = f["data/1"]
    assert g1.attrs["step"] == 3 and g1.attrs["time"] == 3e-3 and g1.attrs["dt"] == 1e-3
    assert "attrs" not in g1                       # attributes, not datasets
    np.testing.assert_array_equal(g1["psi"], np.full(n, 1.0 + 1j))
    assert set(g1["running_state"]) == {"dt", "mu"} and "running_state" not in f["data/0"]
    np.testing.assert_array_equal(f["mesh/edge_mesh/edges"], mesh.edge_mesh.edges)
    np.testing.assert_array_equal(f["mesh/sites"], mesh.sites)
    o = f["solution/options"].attrs
    assert o["save_every"] == 3 and o["solve_time"] == 1.0 and o["sparse_solver"] == "superlu"
    assert "output_file" not in o                  # None values are not attributes (h5py has no null)
    assert f["solution"].attrs["total_seconds"] == 2.0


def test_solution_save_picks_the_backend(tmp_path, caplog):
    """``SolverOptions.output_file`` goes through ``Solution.save``: HDF5 when h5py exists,
    otherwise the npz tree under ``<name>.npz`` with a warning (never a silent change)."""
    import importlib.util
    import logging

    from tdgl_b200.solution import SavedSteps, Solution

    mesh = make_film_mesh(6, 4, 0.5)
    n, E = len(mesh.sites), len(mesh.edge_mesh.edges)
    saved = SavedSteps()
    saved.save_fixed_values({"epsilon": np.ones(n)})
    saved.save_time_step({"step": 0, "time": 0.0, "dt": 1e-3},
                         {"psi": np.ones(n, complex), "mu": np.zeros(n),
                          "supercurrent": np.zeros(E), "normal_current": np.zeros(E)}, None)
    sol = Solution(device=None, options=tdgl.SolverOptions(solve_time=1.0), saved=saved, mesh=mesh)
    assert sol.save(str(tmp_path / "a.npz")) == str(tmp_path / "a.npz")
    if importlib.util.find_spec("h5py") is None:
        with caplog.at_level(logging.WARNING, logger="solver"):
            path = sol.save(str(tmp_path / "b.h5"))
        assert path == str(tmp_path / "b.h5.npz") and os.path.exists(path)
        assert any("h5py is not installed" in r.message for r in caplog.records)
        np.testing.assert_array_equal(Solution.from_npz(path).tdgl_data.psi, np.ones(n))


def test_native_dual_mesh_equals_the_numpy_construction():
    """Mesh.from_triangulation (library: tdgl_host_mesh_dual, several host threads) against
    the vectorised NumPy construction of the same arrays: bit for bit, on a film with a hole
    (enough triangles for several sort chunks and merges), for 1, 3 and 8 host threads; and
    the error behaviour of the native entry point."""
    from tdgl_b200.mesh import Mesh, make_film_points, triangulate

    holes = ((5.0, 3.0, 9.0),)
    pts, tri = triangulate(make_film_points(150, 120, 0.4, holes), holes)
    assert 3 * len(tri) > 4 * (1 << 17)
    ref = Mesh._from_triangulation_numpy(pts, tri)
    for threads in ("1", "3", "8"):
        os.environ["TDGL_B200_HOST_THREADS"] = threads
        try:
            got = Mesh.from_triangulation(pts, tri)
        finally:
            del os.environ["TDGL_B200_HOST_THREADS"]
        for k in ("sites", "elements", "boundary_indices", "areas", "dual_sites"):
            a, b = getattr(got, k), getattr(ref, k)
            assert a.dtype == b.dtype and np.array_equal(a, b), (threads, k)
        for k in ("centers", "edges", "boundary_edge_indices", "directions",
                  "normalized_directions", "edge_lengths", "dual_edge_lengths"):
            a, b = getattr(got.edge_mesh, k), getattr(ref.edge_mesh, k)
            assert a.dtype == b.dtype and np.array_equal(a, b), (threads, k)
    bad = tri.copy()
    bad[7, 1] = len(pts)
    with pytest.raises(ValueError, match="element index out of range"):
        Mesh.from_triangulation(pts, bad)
    # without the dual mesh nothing native is needed (reference create_submesh=False)
    m = Mesh.from_triangulation(pts, tri, create_submesh=False)
    assert m.edge_mesh is None and np.array_equal(m.boundary_indices, ref.boundary_indices)


def test_output_arrays_of_the_step_seam_are_validated():
    """DeviceEngine.update / update_local hand `out` to the library as raw pointers: wrong
    dtype, length or layout is a ValueError, None entries and pinned-style views pass."""
    from tdgl_b200.engine import _check_out

    n, e = 10, 25
    good = (np.empty(n, complex), np.empty(n), np.empty(e), np.empty(e))
    _check_out(good, (n, n, e, e))
    _check_out((None, None, None, None), (n, n, e, e))
    buf = (ctypes.c_char * (16 * n))()                       # what pinned_empty builds on
    view = np.frombuffer(buf, dtype=np.complex128, count=n)
    _check_out((view, good[1], np.empty(e + 3), good[3]), (n, n, e, e))
    for bad in ((np.empty(n), good[1], good[2], good[3]),                    # dtype
                (good[0], np.empty(n - 1), good[2], good[3]),                # too short
                (good[0], good[1], np.empty(2 * e)[::2], good[3]),           # strided
                (good[0], good[1], np.empty((e, 1)), good[3]),               # 2-D
                (good[0], good[1], list(range(e)), good[3]),                 # not an array
                good[:3]):
        with pytest.raises(ValueError):
            _check_out(bad, (n, n, e, e))
