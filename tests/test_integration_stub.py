"""The reference-side ctypes stub of INTEGRATION.md is real code (tools/reference_binding_b200.py)
and is executed here: against the UNMODIFIED reference's TDGLSolver in the build container
(no GPU there: binding, marshalling and the error path), and on the GPU box against a solver
object carrying the reference's attribute names, stepping through ``B200Step.update`` like
``Runner._run_stage`` would (runner.py:417-428) and comparing with the oracle."""
import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tools", "reference_binding_b200.py")


def _load_stub():
    import __graft_entry__ as ge

    ge.build()
    os.environ["TDGL_B200_LIB"] = ge.LIB
    spec = importlib.util.spec_from_file_location("tdgl_b200_reference_binding", STUB)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_integration_md_holds_the_stub_verbatim():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert open(STUB).read() in md


@pytest.mark.reference
def test_stub_binds_against_the_unmodified_reference():
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("/root/reference not present")
    import torch

    from helpers import load_case

    ref = ref_loader.load()
    c = load_case("strip_transport")
    opts = ref.SolverOptions(solve_time=1.0, **{k: v for k, v in c.opts.items() if k != "solve_time"})
    terms = [ref.TerminalInfo(*t) for t in c.terminals]
    solver = ref_loader.make_reference_solver(
        c.mesh, opts, A_applied=c.A, epsilon=c.eps, u=c.u, gamma=c.gamma, terminal_info=terms,
        terminal_currents=c.currents, probe_points=c.probes)
    stub = _load_stub()
    if torch.cuda.is_available():
        step = stub.B200Step(solver, mesh=c.mesh, result_type=ref.solver.SolverResult)
        assert step.h.value
        return
    # no CUDA device in the build container: tdgl_create must fail loudly, and the message
    # must come back as text (restype c_char_p), not as a truncated pointer
    with pytest.raises(RuntimeError, match=r"tdgl_create failed \(3\): .*cuda"):
        stub.B200Step(solver, mesh=c.mesh, result_type=ref.solver.SolverResult)


class _Running:
    """reference RunningState.append (runner.py:214-221)"""

    def __init__(self):
        self.step, self.values = 0, {}

    def append(self, name, value):
        self.values.setdefault(name, []).append(np.array(value))


@pytest.mark.gpu
@pytest.mark.parametrize("dynamic_A", [False, True])
def test_stub_steps_like_the_reference_seam(dynamic_A):
    from helpers import load_case
    from oracle import tdgl_oracle as orc
    from tdgl_b200.solver import SolverResult

    c = load_case("film20_ramp" if dynamic_A else "strip_transport")
    kw = {k: v for k, v in c.opts.items() if k in orc.OracleOptions.__dataclass_fields__}
    cf = (lambda t: c.currents) if c.currents else None

    def oracle():
        return orc.OracleSolver(c.mesh, orc.OracleOptions(**kw), c.A, c.eps, u=c.u, gamma=c.gamma,
                                terminal_info=[orc.TerminalInfo(*t) for t in c.terminals],
                                current_func=cf, probe_points=c.probes,
                                A_func=c.A_func if dynamic_A else None)

    o = oracle()
    # the attribute names B200Step reads from a reference TDGLSolver (solver.py:126-320)
    solver = SimpleNamespace(
        options=o.options,
        operators=SimpleNamespace(fixed_sites=o.fixed_sites), probe_points=c.probes,
        gamma=c.gamma, u=c.u, current_A_applied=np.array(c.A if not dynamic_A else c.A_func(0.0)),
        epsilon=c.eps, mu_boundary=o.mu_boundary, update_mu_boundary=o.update_mu_boundary,
        dynamic_vector_potential=dynamic_A, dynamic_epsilon=False,
        update_applied_vector_potential=(lambda t: np.asarray(c.A_func(t), float)) if dynamic_A else None,
        normalized_directions=o.normalized_directions, device=None)
    stub = _load_stub()
    step = stub.B200Step(solver, mesh=c.mesh, result_type=SolverResult)
    E = len(c.mesh.edge_mesh.edges)
    names = ["psi", "mu", "supercurrent", "normal_current", "induced_vector_potential"]
    values = [o.psi_init.copy(), o.mu_init.copy(), np.zeros(E), np.zeros(E), np.zeros((E, 2))]
    if dynamic_A:
        names.append("applied_vector_potential")
        values.append(solver.current_A_applied)
    running = _Running()
    time, dt, n_steps = 0.0, kw["dt_init"], 60
    for i in range(n_steps):
        res = step.update({"step": i, "time": time, "dt": dt}, running, dt, **dict(zip(names, values)))
        new_dt, *values = res                       # runner.py:424
        values = values[:len(names)]                # (zip(self.names, self.values), runner.py:420)
        dt = new_dt
        time += dt
    ref = orc.run(oracle(), end_time=1e9, max_steps=n_steps)
    got = dict(zip(names, values))
    got["dt"] = np.array([float(v) for v in running.values["dt"]])
    d = orc.compare(got, ref, c.mesh.areas)
    print("stub vs oracle", "dynamic A" if dynamic_A else "transport", d)
    for k, v in d.items():
        assert v < 1e-8, (k, d)
