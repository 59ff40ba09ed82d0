"""Where does the time of the step seam go?  (GPU debugging aid, not a test.)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from tdgl_b200 import SolverOptions, TDGLSolver  # noqa: E402
from tdgl_b200.engine import pinned_empty  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
work = bench.build_workload("film1m_holes_transport")
mesh = work["mesh"]
n, E = len(mesh.sites), len(mesh.edge_mesh.edges)
opts = SolverOptions(solve_time=1e9, save_every=100, **work["opts"])
s = TDGLSolver.from_dimensionless(mesh, opts, A_applied=work["A"], epsilon=work["eps"],
                                  terminal_info=work["terms"], terminal_currents=work["currents"])
eng = s.engine
psi0, mu0 = bench.start_state(work)
eng.set_state(psi0, mu0)
s.update_mu_boundary(0.0)
a = eng.advance(30, 1e300, 0, 0.0)
print("resident ms/step", eng.advance(50, 1e300, a.step, a.time).device_ms / 50)
psi_h, mu_h = pinned_empty(n, np.complex128), pinned_empty(n, np.float64)
out = (pinned_empty(n, np.complex128), pinned_empty(n, np.float64), pinned_empty(E, np.float64),
       pinned_empty(E, np.float64))
p, m = eng.get_state()
psi_h[:] = p
mu_h[:] = m
step, t = 80, 1.0


def loop(label, fn, reps=30):
    fn()
    fn()
    t0 = time.perf_counter()
    dev = 0.0
    for _ in range(reps):
        r = fn()
        dev += r if isinstance(r, float) else 0.0
    el = (time.perf_counter() - t0) / reps * 1e3
    print(f"{label}: {el:.3f} ms/call" + (f" (device-timed part {dev / reps:.3f} ms)" if dev else ""))


def engine_update():
    info, _ = eng.update(psi_h, mu_h, step, t, out=out)
    return float(info.device_ms)


def solver_update():
    res = s.update({"step": step, "time": t, "dt": 1e-2}, None, 1e-2, psi=psi_h, mu=mu_h, out=out)
    return 0.0


def set_state_only():
    eng.set_state(psi_h, mu_h)
    return 0.0


def advance_one():
    return float(eng.advance(1, 1e300, step, t).device_ms)


def get_all():
    eng.get_state()
    eng.get_currents()
    return 0.0


loop("eng.set_state (H2D 24 MB + gathers)", set_state_only)
loop("eng.advance(1) (graph, one step)", advance_one)
loop("eng.update (seam, pinned in/out)", engine_update)
loop("solver.update (seam through TDGLSolver)", solver_update)
loop("get_state + get_currents (pageable)", get_all, reps=10)
