"""GPU debugging aid (round 2): sharded engine after the CG restructure, seeded restart,
terminal_psi=None sensitivity.  Not a test."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np  # noqa: E402

from helpers import load_case  # noqa: E402
from oracle import tdgl_oracle as orc  # noqa: E402


def sharded(world, use_graph, steps=20):
    from tdgl_b200.sharded import LocalShardGroup

    c = load_case("film20_fixed")
    t0 = time.time()
    try:
        grp = LocalShardGroup(c.mesh, world, gamma=c.gamma, u=c.u, probe_sites=c.probes,
                              use_graph=use_graph, running_capacity=1000)
        grp.set_link_exponents(c.A)
        grp.set_epsilon(c.eps)
        o = c.opts
        grp.set_stepper(dt_init=o["dt_init"], dt_max=o["dt_max"], adaptive=o.get("adaptive", True))
        grp.set_state(np.ones(len(c.mesh.sites), complex), np.zeros(len(c.mesh.sites)))
        for k in range(steps):
            info = grp.advance(1, 1e300, k, k * o["dt_init"])
            if k < 3 or k == steps - 1:
                print("  step", k, "its", info.mu_iterations, "res", info.mu_rel_residual, "status", info.status)
        psi, mu = grp.get_state()
        o_ = orc.OracleSolver(c.mesh, orc.OracleOptions(**{k: v for k, v in c.opts.items()
                                                           if k in orc.OracleOptions.__dataclass_fields__}),
                              c.A, c.eps, u=c.u, gamma=c.gamma)
        r = orc.run(o_, end_time=1e9, max_steps=steps)
        print("  sharded", world, "graph" if use_graph == 1 else "host", "vs oracle",
              orc.compare(dict(psi=psi, mu=mu), r, c.mesh.areas), "in %.1fs" % (time.time() - t0))
        grp.close()
    except Exception as exc:
        print("  sharded", world, use_graph, "FAILED after %.1fs:" % (time.time() - t0), repr(exc)[:600])
        traceback.print_exc(limit=3)


def seed_restart():
    from tdgl_b200 import SolverOptions, TDGLSolver

    c = load_case("strip_transport")
    okw = dict(dt_init=c.opts["dt_init"], dt_max=c.opts["dt_max"])

    def make(solve_time, seed=None):
        return TDGLSolver.from_dimensionless(
            c.mesh, SolverOptions(solve_time=solve_time, save_every=50, **okw), A_applied=c.A,
            epsilon=c.eps, terminal_info=c.terminals, terminal_currents=c.currents, u=c.u,
            gamma=c.gamma, seed_solution=seed)

    def oracle(solve_time):
        return orc.OracleSolver(c.mesh, orc.OracleOptions(solve_time=solve_time, **okw), c.A,
                                c.eps, u=c.u, gamma=c.gamma,
                                terminal_info=[orc.TerminalInfo(*t) for t in c.terminals],
                                current_func=lambda t: c.currents)

    first = make(0.8).solve()
    r1 = orc.run(oracle(0.8), end_time=0.8)
    d = first.tdgl_data
    print("  first run vs oracle", orc.compare(dict(psi=d.psi, mu=d.mu, dt=first.dynamics.dt), r1, c.mesh.areas),
          len(first.dynamics.dt), r1["steps"])
    for n in (1, 2, 5, 13, 20):
        s2 = make(0.7, seed=first)
        s2.engine.set_state(d.psi, d.mu)
        s2.update_mu_boundary(0.0)
        info = s2.engine.advance(n, 1e300, 0, 0.0)
        psi, mu = s2.engine.get_state()
        r2 = orc.run(oracle(0.7), end_time=1e9, max_steps=n, psi0=r1["psi"], mu0=r1["mu"])
        print("  seeded, after", n, "steps", orc.compare(dict(psi=psi, mu=mu, dt=s2.engine.get_running(n)[0]), r2, c.mesh.areas))
        # same seed for both: oracle from OUR first-run state
        r3 = orc.run(oracle(0.7), end_time=1e9, max_steps=n, psi0=d.psi, mu0=d.mu)
        print("     oracle seeded with our state:", orc.compare(dict(psi=psi, mu=mu), r3, c.mesh.areas))


def none_sensitivity():
    c = load_case("strip_transport")
    okw = dict(solve_time=1.5, dt_init=c.opts["dt_init"], dt_max=c.opts["dt_max"], terminal_psi=None)

    def oracle():
        return orc.OracleSolver(c.mesh, orc.OracleOptions(**okw), c.A, c.eps, u=c.u, gamma=c.gamma,
                                terminal_info=[orc.TerminalInfo(*t) for t in c.terminals],
                                current_func=lambda t: c.currents)

    r1 = orc.run(oracle(), end_time=1.5)
    o2 = oracle()
    rng = np.random.default_rng(1)
    psi0 = o2.psi_init * (1 + 1e-13 * rng.normal(size=len(c.mesh.sites)))
    r2 = orc.run(o2, end_time=1.5, psi0=psi0)
    print("  terminal_psi=None: oracle vs itself (1e-13 perturbation)", orc.compare(r2, r1, c.mesh.areas))


if __name__ == "__main__":
    import __graft_entry__ as ge

    ge.build()
    which = sys.argv[1:] or ["sharded", "seed", "none"]
    if "sharded" in which:
        for world, ug in ((2, 2), (2, 1), (4, 1)):
            print("sharded", world, ug)
            sharded(world, ug)
    if "seed" in which:
        print("seed restart")
        seed_restart()
    if "none" in which:
        print("terminal_psi None")
        none_sensitivity()
