"""The oracle restatement against the reference's own outputs (committed fixtures made by
oracle/make_golden.py from the unmodified reference).  CPU only."""
import numpy as np
import pytest

from oracle import tdgl_oracle as orc
from helpers import CASES, DYNAMIC_CASES, load_case


def make_oracle(c):
    kw = {k: v for k, v in c.opts.items() if k in orc.OracleOptions.__dataclass_fields__}
    cf = (lambda t: c.currents) if c.currents else None
    return orc.OracleSolver(
        c.mesh, orc.OracleOptions(**kw), c.A, c.eps, u=c.u, gamma=c.gamma,
        terminal_info=[orc.TerminalInfo(*t) for t in c.terminals], current_func=cf,
        probe_points=c.probes, A_func=c.A_func)


@pytest.mark.parametrize("name", CASES)
def test_operator_known_answers(name):
    c = load_case(name)
    g = c.g
    s = make_oracle(c)
    ops = s.operators
    psi, mu, dt = g["op_psi"], g["op_mu"], float(g["op_dt"])
    np.testing.assert_allclose(ops.psi_laplacian @ psi, g["op_lap_psi"], rtol=0, atol=1e-12)
    new_psi, new_sq = orc.solve_for_psi_squared(psi, np.abs(psi) ** 2, mu, s.epsilon,
                                                s.gamma, s.u, dt, ops.psi_laplacian)
    np.testing.assert_allclose(new_psi, g["op_psi_new"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(new_sq, g["op_sq_new"], rtol=0, atol=1e-13)
    js = ops.get_supercurrent(psi)
    np.testing.assert_allclose(js, g["op_supercurrent"], rtol=0, atol=1e-13)
    rhs = ops.divergence @ js - ops.mu_boundary_laplacian @ g["op_mu_boundary"]
    np.testing.assert_allclose(rhs, g["op_rhs"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(ops.mu_laplacian @ mu, g["op_lap_mu"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(-(ops.mu_gradient @ mu), g["op_normal_current"], rtol=0,
                               atol=1e-12)


@pytest.mark.parametrize("name", CASES + DYNAMIC_CASES)
def test_trajectory_matches_reference(name):
    """Same SciPy/SuperLU => the restatement reproduces the reference's raw numbers
    (bit-identical in the build container); 1e-9 leaves room for another BLAS/SuperLU
    build on the GPU box.  The gauge-fixed comparison is the portable statement."""
    c = load_case(name)
    g = c.g
    out = orc.run(make_oracle(c), end_time=c.end_time, max_steps=c.max_steps)
    assert out["steps"] == int(g["steps"])
    np.testing.assert_allclose(out["dt"], g["dt"], rtol=1e-9)
    ref = dict(psi=g["psi"], mu=g["mu"], supercurrent=g["supercurrent"],
               normal_current=g["normal_current"], dt=g["dt"])
    d = orc.compare(out, ref, c.mesh.areas)
    assert d["abs_psi"] < 1e-7, d
    assert d["psi"] < 1e-7, d
    assert d["mu"] < 1e-7, d
    assert d["supercurrent"] < 1e-7 and d["normal_current"] < 1e-7, d
