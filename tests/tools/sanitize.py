"""Small end-to-end run for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tests/tools/sanitize.py"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from helpers import load_case  # noqa: E402
from tdgl_b200 import SolverOptions, TDGLSolver  # noqa: E402
from tdgl_b200.sharded import LocalShardGroup  # noqa: E402

c = load_case("strip_transport")
kw = {k: v for k, v in c.opts.items() if k != "solve_time"}
for graph in (True, False):
    s = TDGLSolver.from_dimensionless(
        c.mesh, SolverOptions(solve_time=0.05, save_every=8, use_cuda_graph=graph, **kw),
        A_applied=c.A, epsilon=c.eps, terminal_info=c.terminals, terminal_currents=c.currents,
        probe_point_indices=c.probes, u=c.u, gamma=c.gamma)
    sol = s.solve()
    print("single", graph, len(sol.dynamics.dt), flush=True)
    del s, sol
n = len(c.mesh.sites)
fixed = np.concatenate([np.asarray(t.site_indices) for t in c.terminals])
with LocalShardGroup(c.mesh, 3, fixed_sites=fixed, fix_psi=True, gamma=c.gamma, u=c.u,
                     probe_sites=c.probes, running_capacity=16, replicate_below=100) as g:
    g.set_link_exponents(c.A)
    g.set_epsilon(c.eps)
    g.set_stepper(dt_init=1e-4, dt_max=1e-1)
    psi0 = np.ones(n, complex)
    psi0[fixed] = 0
    g.set_state(psi0, np.zeros(n))
    info = g.advance(6, 1e300, 0, 0.0)
    g.get_state()
    g.get_currents()
    print("sharded", info.steps_done, flush=True)
