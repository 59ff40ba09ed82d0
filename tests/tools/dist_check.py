"""torchrun worker: film20_fixed on world shards, one process per GPU, through
TDGLSolver(options.distributed=True); rank 0 checks parity with the golden fixture.
  torchrun --nproc-per-node N tests/tools/dist_check.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from helpers import load_case  # noqa: E402
from oracle import tdgl_oracle as orc  # noqa: E402
from tdgl_b200 import SolverOptions, TDGLSolver  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local % torch.cuda.device_count())
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
rank, world = dist.get_rank(), dist.get_world_size()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
c = load_case("film20_fixed")
g = c.g
kw = {k: v for k, v in c.opts.items() if k != "solve_time"}
opts = SolverOptions(solve_time=1e9, save_every=steps, distributed=True,
                     cuda_device=torch.cuda.current_device(), **kw)
s = TDGLSolver.from_dimensionless(c.mesh, opts, A_applied=c.A, epsilon=c.eps,
                                  probe_point_indices=c.probes, u=c.u, gamma=c.gamma)
eng = s.engine
eng.set_state(s.psi_init, s.mu_init)
info = eng.advance(steps, 1e300, 0, 0.0)
psi, mu = eng.get_state()
js, jn = eng.get_currents()
dt, mu_p, th_p = eng.get_running(info.steps_done)
if rank == 0:
    print("shard", eng.shard_info(), info, flush=True)
    if steps == 1000:
        ref = dict(psi=g["psi"], mu=g["mu"], supercurrent=g["supercurrent"],
                   normal_current=g["normal_current"])
        d = orc.compare(dict(psi=psi, mu=mu, supercurrent=js, normal_current=jn), ref,
                        c.mesh.areas)
        print("parity", d, flush=True)
        assert all(v < 1e-8 for v in d.values()), d
        np.testing.assert_allclose(dt, g["dt"], rtol=1e-12)
    print("DIST_OK", world, flush=True)
dist.barrier()
dist.destroy_process_group()
