"""Debug driver for the in-process sharded engine: python tests/tools/shard_check.py [world] [steps] [graph]"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from helpers import load_case  # noqa: E402
from oracle import tdgl_oracle as orc  # noqa: E402
from tdgl_b200.engine import DeviceEngine  # noqa: E402
from tdgl_b200.sharded import LocalShardGroup  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
graph = int(sys.argv[3]) if len(sys.argv) > 3 else 2
c = load_case("film20_fixed")
n = len(c.mesh.sites)


def run(e):
    e.set_link_exponents(c.A)
    e.set_epsilon(c.eps)
    e.set_stepper(dt_init=c.opts["dt_init"], dt_max=c.opts["dt_max"], adaptive=False)
    e.set_state(np.ones(n, complex), np.zeros(n))
    info = e.advance(steps, 1e300, 0, 0.0)
    return info, e.get_state()


with DeviceEngine(c.mesh, gamma=c.gamma, u=c.u, use_graph=graph, running_capacity=steps) as e1:
    i1, (p1, m1) = run(e1)
print("single", i1, flush=True)
devs = os.environ.get("SHARD_DEVICES")
devs = [int(d) for d in devs.split(",")] if devs else None
with LocalShardGroup(c.mesh, world, devices=devs, gamma=c.gamma, u=c.u, use_graph=graph,
                     running_capacity=steps) as grp:
    print("shards", grp.shard_info(), flush=True)
    iw, (pw, mw) = run(grp)
print("sharded", iw, flush=True)
print(f"per step: single {i1.device_ms / steps * 1e3:.0f} us, sharded {iw.device_ms / steps * 1e3:.0f} us;"
      f" {iw.mu_iterations / steps:.1f} CG iterations/step", flush=True)
print("diff", orc.compare(dict(psi=pw, mu=mw), dict(psi=p1, mu=m1), c.mesh.areas))
