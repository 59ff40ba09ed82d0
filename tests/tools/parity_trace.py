"""Lock-step trace of the CUDA engine against the CPU oracle on a golden case: both advance
their own trajectory; every `stride` steps the gauge-fixed differences are printed.
Usage: python tests/tools/parity_trace.py <case|smoke> [steps] [stride] [mu_rtol]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from oracle import tdgl_oracle as orc  # noqa: E402
from tdgl_b200 import SolverOptions, TDGLSolver  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "strip_transport"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 10
mu_rtol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-10

if name == "smoke":
    from tdgl_b200.synthetic import film_problem

    mesh, A, eps, terms = film_problem(12, 12, 0.4, b=0.3, disorder=True)
    kw = dict(dt_init=1e-4, dt_max=1e-1)
    currents, probes, u, gamma = None, None, 5.79, 10.0
    terminals = ()
else:
    from helpers import load_case

    c = load_case(name)
    mesh, A, eps, terminals, currents, probes, u, gamma = (
        c.mesh, c.A, c.eps, c.terminals, c.currents or None, c.probes, c.u, c.gamma)
    kw = {k: v for k, v in c.opts.items() if k != "solve_time"}

okw = {k: v for k, v in kw.items() if k in orc.OracleOptions.__dataclass_fields__}
cf = (lambda t: currents) if currents else None
o = orc.OracleSolver(mesh, orc.OracleOptions(solve_time=1e9, **okw), A, eps, u=u, gamma=gamma,
                     terminal_info=[orc.TerminalInfo(*t) for t in terminals], current_func=cf,
                     probe_points=probes)
opts = SolverOptions(solve_time=1e9, save_every=stride, mu_rtol=mu_rtol, **kw)
s = TDGLSolver.from_dimensionless(mesh, opts, A_applied=A, epsilon=eps, terminal_info=terminals,
                                  terminal_currents=currents, probe_point_indices=probes, u=u,
                                  gamma=gamma)
eng = s.engine
eng.set_state(s.psi_init, s.mu_init)
s.update_mu_boundary(0.0)
psi, mu = o.psi_init.copy(), o.mu_init.copy()
t_o, step, t_e = 0.0, 0, 0.0
a = mesh.areas
while step < steps:
    dts_o = []
    for k in range(stride):
        dt, psi, mu, js, jn = o.update(step + k, t_o, psi, mu)
        t_o += dt
        dts_o.append(dt)
    info = eng.advance(stride, 1e300, step, t_e)
    dts_e = eng.get_running(info.steps_done)[0]
    step, t_e = info.step, info.time
    p, m = eng.get_state()
    js_e, jn_e = eng.get_currents()
    d = orc.compare(dict(psi=p, mu=m, supercurrent=js_e, normal_current=jn_e),
                    dict(psi=psi, mu=mu, supercurrent=js, normal_current=jn), a)
    ddt = float(np.abs(np.array(dts_o) - dts_e).max() / np.abs(dts_o).max())
    print(f"step {step:5d} t {t_o:.4f}/{t_e:.4f} dt {dts_o[-1]:.3e} ddt {ddt:.1e} "
          + " ".join(f"{k}={v:.2e}" for k, v in d.items())
          + f" retries {info.retries} cg {info.mu_iterations} res {info.mu_rel_residual:.1e}",
          flush=True)
