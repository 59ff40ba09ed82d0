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
