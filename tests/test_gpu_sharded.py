"""The domain-decomposed engine on the GPU.  Several shards of one mesh run
in this process, one host thread each, on the same device: the halo-exchange and all-reduce
kernels execute exactly as across GPUs (stores / flag spins on the peers' arenas), so a
1-GPU box covers the sharded code path; with >= 2 GPUs the same tests also run one shard per
device, and a torchrun job exercises the one-process-per-GPU wiring (CUDA IPC)."""
import os
import subprocess
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # one hardware queue per shard stream

import numpy as np  # noqa: E402
import pytest  # noqa: E402

from helpers import load_case  # noqa: E402
from oracle import tdgl_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _devices(world):
    import torch

    n = torch.cuda.device_count()
    return [r % n for r in range(world)] if n >= world else [0] * world


def _group(c, world, devices=None, **kw):
    from tdgl_b200.sharded import LocalShardGroup

    fixed = (np.concatenate([np.asarray(t.site_indices) for t in c.terminals])
             if c.terminals else None)
    g = LocalShardGroup(c.mesh, world, devices=devices, fixed_sites=fixed, fix_psi=True,
                        gamma=c.gamma, u=c.u, probe_sites=c.probes, **kw)
    g.set_link_exponents(c.A)
    g.set_epsilon(c.eps)
    return g


def _stepper(c):
    o = c.opts
    return dict(dt_init=o["dt_init"], dt_max=o["dt_max"], adaptive=o.get("adaptive", True))


@pytest.mark.parametrize("world,use_graph,replicate_below",
                         [(2, 1, 0), (4, 1, 1), (3, 2, 100), (8, 1, 0), (2, 2, 0)])
def test_sharded_smooth_trajectory_matches_reference(world, use_graph, replicate_below):
    """film20_fixed, 1000 fixed-dt steps on `world` shards: same 1e-8 gauge-fixed parity with
    the reference as the single-GPU engine.  replicate_below = 1 partitions every AMG level
    but the coarsest, 100 all but the last two, 0 (default) only the fine level here."""
    c = load_case("film20_fixed")
    g = c.g
    with _group(c, world, use_graph=use_graph, running_capacity=1000,
                replicate_below=replicate_below) as grp:
        grp.set_stepper(**_stepper(c))
        grp.set_state(np.ones(len(c.mesh.sites), complex), np.zeros(len(c.mesh.sites)))
        info = grp.advance(1000, 1e300, 0, 0.0)
        assert info.steps_done == 1000
        psi, mu = grp.get_state()
        js, jn = grp.get_currents()
        dt, mu_p, th_p = grp.get_running(1000)
        print("shards", grp.shard_info(), "info", grp.info(), info)
        assert grp.info()["graph_mode"] == use_graph
    ref = dict(psi=g["psi"], mu=g["mu"], supercurrent=g["supercurrent"],
               normal_current=g["normal_current"])
    d = orc.compare(dict(psi=psi, mu=mu, supercurrent=js, normal_current=jn), ref, c.mesh.areas)
    print("film20_fixed on", world, "shards:", d)
    for k, v in d.items():
        assert v < 1e-8, (k, d)
    np.testing.assert_allclose(dt, g["dt"], rtol=1e-12)
    # probe traces: gauge-invariant combination (voltage between the two probes)
    v_ref = g["running_mu"][0] - g["running_mu"][1]
    np.testing.assert_allclose(mu_p[0] - mu_p[1], v_ref, atol=1e-8 * np.abs(v_ref).max())


def test_sharded_transport_matches_reference():
    """Terminals (fixed sites, boundary currents), holes, adaptive dt with retries, probes —
    on 3 shards; parity on the smooth window (see test_gpu_parity.test_transport_trajectory)."""
    c = load_case("strip_transport")
    g = c.g
    n = len(c.mesh.sites)
    with _group(c, 3, running_capacity=150) as grp:
        grp.set_stepper(**_stepper(c))
        psi0 = np.ones(n, complex)
        fixed = np.concatenate([np.asarray(t.site_indices) for t in c.terminals])
        psi0[fixed] = 0.0
        grp.set_state(psi0, np.zeros(n))
        mub = np.zeros(len(c.mesh.edge_mesh.boundary_edge_indices))
        names = [t.name for t in c.terminals]
        for t in c.terminals:
            dens = (-1 / t.length) * sum(c.currents[m] for m in names if m != t.name)
            mub[np.asarray(t.boundary_edge_indices)] = dens
        grp.set_mu_boundary(mub)
        info = grp.advance(150, 1e300, 0, 0.0)
        psi, mu = grp.get_state()
        dt = grp.get_running(150)[0]
    k = list(g["snap_steps"]).index(150)
    d = orc.compare(dict(psi=psi, mu=mu), dict(psi=g["snap_psi"][k], mu=g["snap_mu"][k]),
                    c.mesh.areas)
    print("strip_transport on 3 shards, step 150:", d, info)
    assert d["psi"] < 1e-8 and d["mu"] < 1e-8, d
    np.testing.assert_allclose(dt, g["dt"][:150], rtol=1e-8)
    assert np.abs(psi[fixed]).max() == 0.0


def test_sharded_large_mesh_against_single_engine():
    """250k sites on 4 shards against the single-shard engine: identical step bookkeeping,
    psi / mu within 1e-9 after 30 adaptive steps (summation orders differ)."""
    from tdgl_b200.engine import DeviceEngine
    from tdgl_b200.sharded import LocalShardGroup
    from tdgl_b200.synthetic import film_problem

    mesh, A, eps, _ = film_problem(200, 200, 0.43, b=0.1)
    n = len(mesh.sites)
    st = dict(dt_init=1e-4, dt_max=1e-1)

    def run(e):
        e.set_link_exponents(A)
        e.set_epsilon(eps)
        e.set_stepper(**st)
        e.set_state(np.ones(n, complex), np.zeros(n))
        info = e.advance(30, 1e300, 0, 0.0)
        return info, e.get_state(), e.get_currents(), e.get_running(30)[0]

    with DeviceEngine(mesh, running_capacity=64) as e1:
        i1, (p1, m1), (js1, jn1), dt1 = run(e1)
    with LocalShardGroup(mesh, 4, devices=_devices(4), running_capacity=64) as grp:
        i4, (p4, m4), (js4, jn4), dt4 = run(grp)
        print("250k on 4 shards:", grp.shard_info())
    assert (i1.step, i1.retries) == (i4.step, i4.retries)
    np.testing.assert_allclose(dt4, dt1, rtol=1e-10)
    d = orc.compare(dict(psi=p4, mu=m4, supercurrent=js4, normal_current=jn4),
                    dict(psi=p1, mu=m1, supercurrent=js1, normal_current=jn1), mesh.areas)
    print("250k, 4 shards vs 1:", d, "iterations", i1.mu_iterations, i4.mu_iterations)
    for k, v in d.items():
        assert v < 1e-8, (k, d)
    assert abs(i4.mu_iterations - i1.mu_iterations) <= 0.1 * i1.mu_iterations + 5


def test_one_process_per_gpu_torchrun():
    """torchrun, NCCL for the plumbing, CUDA IPC for the peer arenas (needs >= 2 GPUs)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "tools", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_OK" in res.stdout


def test_runs_are_bitwise_reproducible():
    """Reductions are added in a fixed order (blocks, then ranks): like the reference, two
    runs of the same problem give identical bits — single engine and 3 shards."""
    from tdgl_b200.engine import DeviceEngine

    c = load_case("film20_adaptive")
    n = len(c.mesh.sites)

    def run(e):
        e.set_link_exponents(c.A)
        e.set_epsilon(c.eps)
        e.set_stepper(**_stepper(c))
        e.set_state(np.ones(n, complex), np.zeros(n))
        info = e.advance(60, 1e300, 0, 0.0)
        return info[:7] + info[8:12], e.get_state(), e.get_currents()

    def same(a, b):
        assert a[0] == b[0]
        for x, y in zip(a[1] + a[2], b[1] + b[2]):
            np.testing.assert_array_equal(x, y)

    with DeviceEngine(c.mesh, gamma=c.gamma, u=c.u, running_capacity=64) as e:
        r1 = run(e)
    with DeviceEngine(c.mesh, gamma=c.gamma, u=c.u, running_capacity=64) as e:
        same(r1, run(e))
    with _group(c, 3, running_capacity=64) as g:
        s1 = run(g)
    with _group(c, 3, running_capacity=64) as g:
        same(s1, run(g))


@pytest.mark.parametrize("world", [1, 3])
def test_shard_local_step_seam_matches_reference(world):
    """tdgl_update_local: every shard is handed, and hands back, only the entries it owns (no
    whole-mesh arrays, no sums over zero-padded outputs); halo values travel between the
    devices.  40 steps threaded through the seam on the transport strip (terminals, adaptive
    dt) against the golden trajectory of the reference at 1e-8."""
    c = load_case("strip_transport")
    g = c.g
    n, E = len(c.mesh.sites), len(c.mesh.edge_mesh.edges)
    with _group(c, world, running_capacity=64) as grp:
        grp.set_stepper(**_stepper(c))
        fixed = np.concatenate([np.asarray(t.site_indices) for t in c.terminals])
        psi0 = np.ones(n, complex)
        psi0[fixed] = 0.0
        mub = np.zeros(len(c.mesh.edge_mesh.boundary_edge_indices))
        names = [t.name for t in c.terminals]
        for t in c.terminals:
            dens = (-1 / t.length) * sum(c.currents[m] for m in names if m != t.name)
            mub[np.asarray(t.boundary_edge_indices)] = dens
        grp.set_mu_boundary(mub)
        grp.set_state(psi0, np.zeros(n))       # (only so that every mailbox starts valid)
        maps = grp.local_maps()
        assert sorted(np.concatenate([m[0] for m in maps])) == list(range(n))
        assert sorted(np.concatenate([m[1] for m in maps])) == list(range(E))
        psi = [psi0[m[0]].copy() for m in maps]
        mu = [np.zeros(len(m[0])) for m in maps]
        outs = [(np.empty(len(m[0]), complex), np.empty(len(m[0])), np.empty(len(m[1])),
                 np.empty(len(m[1]))) for m in maps]
        step, time, dts = 0, 0.0, []
        for _ in range(40):
            info, outs = grp.update_local(psi, mu, step, time, outs)
            psi = [o[0].copy() for o in outs]
            mu = [o[1].copy() for o in outs]
            dts.append(info.dt)
            step, time = info.step, info.time
        full_psi, full_mu = np.zeros(n, complex), np.zeros(n)
        js, jn = np.zeros(E), np.zeros(E)
        for m, o in zip(maps, outs):
            full_psi[m[0]], full_mu[m[0]] = o[0], o[1]
            js[m[1]], jn[m[1]] = o[2], o[3]
    o = orc.OracleSolver(c.mesh, orc.OracleOptions(solve_time=1e9, dt_init=c.opts["dt_init"],
                                                   dt_max=c.opts["dt_max"]), c.A, c.eps, u=c.u,
                         gamma=c.gamma, terminal_info=[orc.TerminalInfo(*t) for t in c.terminals],
                         current_func=lambda t: c.currents)
    ref = orc.run(o, end_time=1e9, max_steps=40)
    d = orc.compare(dict(psi=full_psi, mu=full_mu, supercurrent=js, normal_current=jn,
                         dt=np.array(dts)), ref, c.mesh.areas)
    print("shard-local seam on", world, "shards, 40 steps:", d)
    for k, v in d.items():
        assert v < 1e-8, (k, d)
    del g


@pytest.mark.parametrize("world", [1, 2])
def test_shard_local_seam_fed_back_equals_the_resident_loop(world):
    """A Runner-style loop through tdgl_update_local that feeds every step's output back as
    the next input must be the device-resident loop, bit for bit: the engine recognises the
    state it already holds (bitwise compare + a vote of all ranks), replaces nothing and keeps
    the initial-guess history of the mu solve.  A state that differs in one bit takes the other
    path (state replaced, history reset) and still agrees to solver accuracy."""
    c = load_case("strip_transport")
    n, E = len(c.mesh.sites), len(c.mesh.edge_mesh.edges)
    fixed = np.concatenate([np.asarray(t.site_indices) for t in c.terminals])
    psi0 = np.ones(n, complex)
    psi0[fixed] = 0.0
    mub = np.zeros(len(c.mesh.edge_mesh.boundary_edge_indices))
    names = [t.name for t in c.terminals]
    for t in c.terminals:
        dens = (-1 / t.length) * sum(c.currents[m] for m in names if m != t.name)
        mub[np.asarray(t.boundary_edge_indices)] = dens
    steps = 24

    def prepare(grp):
        grp.set_stepper(**_stepper(c))
        grp.set_mu_boundary(mub)
        grp.set_state(psi0, np.zeros(n))

    with _group(c, world, running_capacity=64) as grp:
        prepare(grp)
        a = grp.advance(steps, 1e300, 0, 0.0)
        psi_loop, mu_loop = grp.get_state()

    def seam(perturb_at=None):
        with _group(c, world, running_capacity=64) as grp:
            prepare(grp)
            maps = grp.local_maps()
            psi = [psi0[m[0]].copy() for m in maps]
            mu = [np.zeros(len(m[0])) for m in maps]
            outs = [(np.empty(len(m[0]), complex), np.empty(len(m[0])), np.empty(len(m[1])),
                     np.empty(len(m[1]))) for m in maps]
            step, time, its = 0, 0.0, 0
            for k in range(steps):
                if k == perturb_at:     # one bit of one entry on one shard
                    v = mu[-1][:1].view(np.uint64)
                    v ^= np.uint64(1)
                info, outs = grp.update_local(psi, mu, step, time, outs)
                psi = [o[0].copy() for o in outs]
                mu = [o[1].copy() for o in outs]
                step, time = info.step, info.time
                its += info.mu_iterations
            full_psi, full_mu = np.zeros(n, complex), np.zeros(n)
            for m, o in zip(maps, outs):
                full_psi[m[0]], full_mu[m[0]] = o[0], o[1]
        return full_psi, full_mu, its, time

    psi_s, mu_s, its_s, t_s = seam()
    assert t_s == a.time and its_s == a.mu_iterations, (t_s, a.time, its_s, a.mu_iterations)
    assert np.array_equal(psi_s, psi_loop) and np.array_equal(mu_s, mu_loop)
    psi_p, mu_p, its_p, _ = seam(perturb_at=steps // 2)
    d = orc.compare(dict(psi=psi_p, mu=mu_p), dict(psi=psi_loop, mu=mu_loop), c.mesh.areas)
    print("seam with one flipped bit at step", steps // 2, "vs the loop:", d, "mu iterations",
          its_p, "vs", its_s)
    assert d["psi"] < 1e-8 and d["mu"] < 1e-8, d
    assert its_p >= its_s - 2   # (the reset guess costs iterations, it does not save them)
