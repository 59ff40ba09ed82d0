"""Small stepping run for compute-sanitizer (host-driven launches: every kernel is an ordinary
launch the tools can instrument).  argv[1]: number of shards (1 = single engine)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np  # noqa: E402

from tdgl_b200.synthetic import film_problem  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mesh, A, eps, terms = film_problem(10, 6, 0.5, b=0.3, terminals=True)
n = len(mesh.sites)
fixed = np.concatenate([np.asarray(t.site_indices) for t in terms])
kw = dict(fixed_sites=fixed, fix_psi=True, use_graph=2, running_capacity=16)
if world == 1:
    from tdgl_b200.engine import DeviceEngine

    eng = DeviceEngine(mesh, **kw)
else:
    from tdgl_b200.sharded import LocalShardGroup

    eng = LocalShardGroup(mesh, world, **kw)
eng.set_link_exponents(A)
eng.set_epsilon(eps)
eng.set_stepper(dt_init=1e-3, dt_max=1e-2, adaptive=True)
psi0 = np.ones(n, complex)
psi0[fixed] = 0
eng.set_state(psi0, np.zeros(n))
mub = np.zeros(len(mesh.edge_mesh.boundary_edge_indices))
for t, dens in zip(terms, (0.1, -0.1)):
    mub[np.asarray(t.boundary_edge_indices)] = dens
eng.set_mu_boundary(mub)
info = eng.advance(steps, 1e300, 0, 0.0)
psi, mu = eng.get_state()
js, jn = eng.get_currents()
print("sanitize target:", world, "shard(s),", info.steps_done, "steps, mu iterations", info.mu_iterations,
      "|psi| in", float(np.abs(psi).min()), float(np.abs(psi).max()), "sites", n)
eng.close()
