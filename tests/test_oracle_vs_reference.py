"""The oracle restatement against the unmodified reference executed live (only where
/root/reference exists, i.e. the build container)."""
import numpy as np
import pytest

from oracle import ref_loader as rl
from oracle import tdgl_oracle as orc
from tdgl_b200.mesh import Mesh
from tdgl_b200.synthetic import film_problem

pytestmark = pytest.mark.skipif(not rl.available(), reason="needs /root/reference")


def test_mesh_arrays_match_reference():
    mesh, *_ = film_problem(12, 8, 0.4, holes=((2.0, 0.5, 1.5),), reorder=False)
    rm = rl.make_reference_mesh(mesh.sites, mesh.elements)
    assert np.array_equal(mesh.edge_mesh.edges, rm.edge_mesh.edges)
    assert np.array_equal(mesh.boundary_indices, rm.boundary_indices)
    assert np.array_equal(mesh.edge_mesh.boundary_edge_indices,
                          rm.edge_mesh.boundary_edge_indices)
    for k in ("centers", "directions", "edge_lengths", "dual_edge_lengths"):
        np.testing.assert_allclose(getattr(mesh.edge_mesh, k), getattr(rm.edge_mesh, k),
                                   rtol=0, atol=1e-14)
    np.testing.assert_allclose(mesh.areas, rm.areas, rtol=1e-12)


def test_mesh_smooth_matches_reference():
    """Mesh.smooth against the unmodified reference's (finite_volume/mesh.py:245-283): the
    same vertex positions after 1 and 5 Laplacian sweeps, boundary vertices untouched."""
    mesh, *_ = film_problem(12, 8, 0.4, holes=((2.0, 0.5, 1.5),), reorder=False)
    rm = rl.make_reference_mesh(mesh.sites, mesh.elements)
    for it in (1, 5):
        ours = mesh.smooth(it)
        theirs = rm.smooth(it)
        np.testing.assert_allclose(ours.sites, theirs.sites, rtol=0, atol=1e-13)
        assert np.array_equal(ours.elements, theirs.elements)
        np.testing.assert_allclose(ours.areas, theirs.areas, rtol=1e-11)
        b = mesh.boundary_indices
        assert np.array_equal(ours.sites[b], mesh.sites[b])
    assert np.abs(mesh.smooth(5).sites - mesh.sites).max() > 1e-3    # it did move


@pytest.mark.parametrize("terminals", [False, True])
def test_oracle_reproduces_reference(terminals):
    ref = rl.load()
    mesh, A, eps, terms = film_problem(16, 8, 0.5, b=0.2, disorder=True,
                                       terminals=terminals)
    cur = {"source": 1.5, "drain": -1.5} if terminals else None
    okw = dict(solve_time=3.0, dt_init=1e-4, dt_max=1e-1)
    rs = rl.make_reference_solver(
        mesh, ref.SolverOptions(**okw), A_applied=A, epsilon=eps,
        terminal_info=[ref.TerminalInfo(*t) for t in terms], terminal_currents=cur)
    r = rl.run_reference(rs, end_time=3.0)
    os_ = orc.OracleSolver(mesh, orc.OracleOptions(**okw), A, eps,
                           terminal_info=[orc.TerminalInfo(*t) for t in terms],
                           current_func=(lambda t: cur) if cur else None)
    o = orc.run(os_, end_time=3.0)
    assert o["steps"] == r["steps"]
    for k in ("psi", "mu", "supercurrent", "normal_current", "dt"):
        np.testing.assert_allclose(o[k], r[k], rtol=0, atol=1e-12)


def _two_stage_reference(rs, skip_time, solve_time):
    """Runner.run (runner.py:288-328) with the reference's own update(): thermalise, then the
    saved stage from step 0 / time 0 with the solver's controller state carried over."""
    th = rl.run_reference(rs, end_time=skip_time)
    return rl.run_reference(rs, end_time=solve_time, psi0=th["psi"], mu0=th["mu"])


@pytest.mark.parametrize("terminal_psi", [0.0, 1.0, None])
def test_oracle_edge_cases_of_the_reference_tests(terminal_psi):
    """What tdgl/test/test_solve.py::test_source_drain_current exercises on this path:
    terminal_psi in {0, 1, None}, a callable terminal current, a time-dependent epsilon
    (``disorder_epsilon(r, *, t)``, solver.py:364-381) and a thermalisation stage."""
    ref = rl.load()
    mesh, A, eps, terms = film_problem(16, 8, 0.5, b=0.2, disorder=True, terminals=True)

    def cur(t):
        return {"source": 1.0 + 0.5 * min(t, 1.0), "drain": -(1.0 + 0.5 * min(t, 1.0))}

    def eps_t(t):
        return eps * (1.0 - 0.2 * min(t / 2.0, 1.0))

    okw = dict(solve_time=2.0, skip_time=0.5, dt_init=1e-4, dt_max=1e-1,
               terminal_psi=terminal_psi)
    rs = rl.make_reference_solver(
        mesh, ref.SolverOptions(**okw), A_applied=A, epsilon=eps_t(0.0),
        terminal_info=[ref.TerminalInfo(*t) for t in terms], current_func=cur)
    rs.dynamic_epsilon = True
    rs.update_epsilon = eps_t
    # (run_reference threads `epsilon` through update() like Runner does when it is dynamic)
    r = _two_stage_reference(rs, 0.5, 2.0)
    os_ = orc.OracleSolver(mesh, orc.OracleOptions(**okw), A, eps_t(0.0),
                           terminal_info=[orc.TerminalInfo(*t) for t in terms],
                           current_func=cur, epsilon_func=eps_t)
    o = orc.run_stages(os_)
    assert o["steps"] == r["steps"]
    for k in ("psi", "mu", "supercurrent", "normal_current", "dt"):
        np.testing.assert_allclose(o[k], r[k], rtol=0, atol=1e-12)
    fixed = np.concatenate([np.asarray(t.site_indices) for t in terms])
    # (identity rows keep psi = 0 exactly; a terminal_psi of 1 only sets the initial value —
    # the reference's fixed rows still evolve through the local terms of the update)
    if terminal_psi == 0.0:
        assert np.abs(o["psi"][fixed]).max() == 0.0


def test_oracle_screening_matches_reference():
    """Row S: the Polyak iteration on the induced vector potential (solver.py:522-578,
    650-688) with the reference's own numba kernel (screening.py:12-42) and
    ``Mesh.get_quantity_on_site`` (mesh.py:203-243)."""
    from types import SimpleNamespace

    ref = rl.load()
    mesh, A, eps, _ = film_problem(10, 6, 0.5, b=0.4, reorder=False)
    rmesh = rl.make_reference_mesh(mesh.sites, mesh.elements)
    scale = 0.2
    okw = dict(solve_time=0.3, dt_init=1e-4, dt_max=1e-2, include_screening=True,
               screening_tolerance=1e-3, max_iterations_per_step=1000)
    rs = rl.make_reference_solver(mesh, ref.SolverOptions(**okw), A_applied=A, epsilon=eps)
    rs.device = SimpleNamespace(mesh=rmesh)
    rs.areas = scale * np.asarray(mesh.areas, float)          # solver.py:309
    rs.sites = np.asarray(mesh.sites, float)
    rs.edge_centers = np.asarray(mesh.edge_mesh.centers, float)
    rs.new_A_induced = np.empty((len(mesh.edge_mesh.edges), 2))
    r = rl.run_reference(rs, end_time=0.3)
    os_ = orc.OracleSolver(mesh, orc.OracleOptions(**okw), A, eps, screening_scale=scale)
    o = orc.run(os_, end_time=0.3)
    assert o["steps"] == r["steps"]
    # the numba kernel (fastmath, parallel) and the restatement differ in the last bit of
    # A_induced; from there SuperLU's null-space component (the gauge) differs: compare
    # gauge-fixed, as everywhere two solvers are compared
    d = orc.compare(o, r, mesh.areas)
    for k, v in d.items():
        assert v < 1e-9, (k, d)
    np.testing.assert_allclose(os_.A_induced, r["induced_vector_potential"], rtol=0, atol=1e-11)
    np.testing.assert_array_equal(o["screening_iterations"],
                                  r["running"]["screening_iterations"][0])
    assert np.abs(os_.A_induced).max() > 1e-4 and o["screening_iterations"].max() > 2
