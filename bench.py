#!/usr/bin/env python
"""Benchmark of the TDGL per-step hot path (BASELINE.json metric: TDGL time-steps/sec and
mesh-sites x steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload film1m_holes_transport|film250k_field|film20_cpu]

One bench "step" = one TDGL time step (one ``TDGLSolver.update``): fused psi update, right
hand side, mu solve, step controller.  Default workload = BASELINE.json configs[2], the
configuration the metric's 1-GPU target is quoted on: ~1.0M-site square film with four
circular holes, source/drain terminals and a transport current, adaptive dt.

b200 arm
  value   whole-job steps/s with the state resident in HBM, timed with CUDA events on the
          engine's stream (tdgl_advance_info.device_ms), barrier + synchronize on both
          sides, max over ranks.
  e2e     the same steps taken one by one through ``TDGLSolver.update`` — the reference's
          own step seam (tdgl/solver/runner.py:417-423) — with host psi/mu in pinned
          memory copied to the device and psi', mu', J_s, J_n copied back EVERY step.
  roofline  dominant kernel of the step, timed on its own with CUDA events after an L2
          flush; algorithmic bytes per DESIGN.md.  roofline.in_loop: the same kernels'
          durations INSIDE the stepping CUDA graph (%globaltimer, tdgl_get_trace).
  cpu_baseline  the oracle port of the reference's scipy.sparse/SuperLU step on this
          box's host cores (N=1 only; bounded number of steps of the same workload).
reference arm (--impl reference): that CPU path alone, same workload/metric.

The metric is BASELINE.json's "TDGL time-steps/sec (and mesh-sites x steps/sec)": `value` is
site-steps/s (sites of the mesh x steps/s), which is comparable across mesh sizes; steps/s is
reported next to it (`steps_per_sec`).

With N > 1 ranks (torchrun) the mesh is domain-decomposed over the N GPUs (one shard per
rank; halo exchange + all-reduces are device code on NVLink peer memory):
  --mode weak (default)  the workload is BASELINE.json's configuration for N GPUs — ~1M sites
                         per GPU: 2M film + holes + transport at N=2, 4M transport strip at
                         N=4, 10M film in a field at N=8 — "scaling": "weak";
  --mode strong          the N=1 workload (or --workload) on N GPUs, "scaling": "strong";
  --mode replicas        every rank steps its own replica of the N=1 workload (no exchange).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

WORKLOADS = {
    # BASELINE.json configs[2]
    "film1m_holes_transport": dict(
        width=400.0, height=400.0, h=0.4225, b=0.0,
        holes=((100.0, 100.0, 20.0), (-100.0, 100.0, 20.0), (100.0, -100.0, 20.0),
               (-100.0, -100.0, 20.0)),
        terminals=True, current=80.0, opts=dict(dt_init=1e-4, dt_max=1e-1, adaptive=True)),
    # BASELINE.json configs[1]
    "film250k_field": dict(
        width=200.0, height=200.0, h=0.43, b=0.1, holes=(), terminals=False, current=0.0,
        opts=dict(dt_init=1e-4, dt_max=1e-1, adaptive=True)),
    # the configs[2] film at twice the area (2 GPUs)
    "film2m_holes_transport": dict(
        width=566.0, height=566.0, h=0.4225, b=0.0,
        holes=((141.5, 141.5, 28.3), (-141.5, 141.5, 28.3), (141.5, -141.5, 28.3),
               (-141.5, -141.5, 28.3)),
        terminals=True, current=113.2, opts=dict(dt_init=1e-4, dt_max=1e-1, adaptive=True)),
    # the same film at 4x and 8x the area: the weak-scaling series is ONE geometry family
    # (holes film, ~1M sites per GPU, same current density)
    "film4m_holes_transport": dict(
        width=800.0, height=800.0, h=0.4225, b=0.0,
        holes=((200.0, 200.0, 40.0), (-200.0, 200.0, 40.0), (200.0, -200.0, 40.0),
               (-200.0, -200.0, 40.0)),
        terminals=True, current=160.0, opts=dict(dt_init=1e-4, dt_max=1e-1, adaptive=True)),
    "film8m_holes_transport": dict(
        width=1131.4, height=1131.4, h=0.4225, b=0.0,
        holes=((282.85, 282.85, 56.57), (-282.85, 282.85, 56.57), (282.85, -282.85, 56.57),
               (-282.85, -282.85, 56.57)),
        terminals=True, current=226.28, opts=dict(dt_init=1e-4, dt_max=1e-1, adaptive=True)),
    # BASELINE.json configs[3]: long strip with transport current (4 GPUs)
    "strip4m_transport": dict(
        width=3200.0, height=200.0, h=0.4225, b=0.0, holes=(), terminals=True, current=40.0,
        opts=dict(dt_init=1e-4, dt_max=1e-1, adaptive=True)),
    # BASELINE.json configs[4]: 10M-site film in a field (strong scaling at 1/2/4/8 GPUs)
    "film10m_field": dict(
        width=1265.0, height=1265.0, h=0.4225, b=0.1, holes=(), terminals=False, current=0.0,
        opts=dict(dt_init=1e-4, dt_max=1e-1, adaptive=True)),
    "film4m_field": dict(
        width=800.0, height=800.0, h=0.4225, b=0.1, holes=(), terminals=False, current=0.0,
        opts=dict(dt_init=1e-4, dt_max=1e-1, adaptive=True)),
    # BASELINE.json configs[0] geometry (perturbed so that every term is active)
    "film20_cpu": dict(
        width=20.0, height=20.0, h=0.29, b=0.3, holes=(), terminals=False, current=0.0,
        disorder=True, opts=dict(dt_init=1e-3, dt_max=1e-3, adaptive=False)),
}


def _mesh_to_arrays(mesh):
    em = mesh.edge_mesh
    return dict(sites=mesh.sites, elements=mesh.elements, boundary_indices=mesh.boundary_indices,
                areas=mesh.areas, centers=em.centers, edges=em.edges,
                boundary_edge_indices=em.boundary_edge_indices, directions=em.directions,
                edge_lengths=em.edge_lengths, dual_edge_lengths=em.dual_edge_lengths)


def _mesh_from_arrays(g):
    from tdgl_b200.mesh import EdgeMesh, Mesh

    em = EdgeMesh(g["centers"], g["edges"], g["boundary_edge_indices"], g["directions"],
                  g["edge_lengths"], g["dual_edge_lengths"])
    return Mesh(g["sites"], g["elements"], g["boundary_indices"], areas=g["areas"], edge_mesh=em)


def auto_workload(n_gpus: int, mode: str) -> str:
    """BASELINE.json's configuration for this GPU count (weak mode) or the 1-GPU one."""
    if mode != "weak" or n_gpus == 1:
        return "film1m_holes_transport"
    return {2: "film2m_holes_transport", 4: "film4m_holes_transport",
            8: "film8m_holes_transport"}.get(n_gpus, "film4m_holes_transport")


START = ("analytic developed state (tdgl_b200.synthetic.vortex_state: vortex/antivortex pairs at"
         " the holes, edge vortices, transport phase gradient), identical in both arms,"
         " then --warmup untimed steps")


def start_state(work):
    """The state both arms start from (see START)."""
    from tdgl_b200.synthetic import vortex_state

    w = WORKLOADS[work["name"]]
    fixed = (np.concatenate([np.asarray(t.site_indices) for t in work["terms"]])
             if work["terms"] else None)
    q = 0.75 * w["current"] / w["width"] if w["terminals"] else 0.0   # J = |psi|^2 q
    return vortex_state(work["mesh"], w["holes"], fixed, q=q, b=w["b"])


def run_config(work):
    """The `config` object: the same keys and values in both arms."""
    return {"workload": work["name"], "sites": len(work["mesh"].sites),
            "edges": len(work["mesh"].edge_mesh.edges), "start": START}


def build_workload(name: str, rank: int = 0, barrier=None):
    """Synthetic mesh + inputs.  With several ranks, rank 0 builds the mesh once and the
    others read it from /dev/shm (the Delaunay triangulation is the slow, single-threaded
    part of the setup and is not what is measured)."""
    from tdgl_b200.synthetic import box_terminal, gaussian_disorder, film_problem, \
        uniform_field_vector_potential

    w = WORKLOADS[name]
    t0 = time.perf_counter()
    if barrier is None:
        mesh, A, eps, terms = film_problem(w["width"], w["height"], w["h"], b=w["b"],
                                           holes=w["holes"], terminals=w["terminals"],
                                           disorder=w.get("disorder", False))
    else:
        shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
        path = os.path.join(shm, f"tdgl_b200_{name}_{os.environ.get('MASTER_PORT', '0')}.npz")
        if rank == 0:
            mesh = film_problem(w["width"], w["height"], w["h"], b=w["b"], holes=w["holes"],
                                terminals=False)[0]
            np.savez(path + ".tmp.npz", **_mesh_to_arrays(mesh))
            os.replace(path + ".tmp.npz", path)
        barrier()
        if rank != 0:
            with np.load(path) as g:
                mesh = _mesh_from_arrays({k: g[k] for k in g.files})
        barrier()
        if rank == 0:
            os.remove(path)
        A = uniform_field_vector_potential(mesh.edge_mesh.centers, w["b"])
        eps = (gaussian_disorder(mesh.sites) if w.get("disorder", False)
               else np.ones(len(mesh.sites)))
        terms = ()
        if w["terminals"]:
            tol = 1e-9 * max(w["width"], w["height"])
            x0, x1, big = -w["width"] / 2, w["width"] / 2, 10 * max(w["width"], w["height"])
            terms = tuple(sorted((box_terminal(mesh, "source", x0 - tol, x0 + tol, -big, big),
                                  box_terminal(mesh, "drain", x1 - tol, x1 + tol, -big, big)),
                                 key=lambda t: t.length))
    currents = None
    if w["terminals"]:
        currents = {"source": w["current"], "drain": -w["current"]}
    return dict(name=name, mesh=mesh, A=A, eps=eps, terms=terms, currents=currents,
                opts=w["opts"], mesh_seconds=time.perf_counter() - t0)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        for key in ("hbm_gbs", "hbm_GBps", "hbm_gb_s"):
            if key in d:
                return float(d[key]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the device is under load.
    `start()` before the warm-up steps, `stop()` after the end-to-end loop: the timed region of
    a default run is only tens of milliseconds long (shorter than nvidia-smi's start-up), so the
    samples cover warm-up + timed steps + the end-to-end steps, all of them stepping load."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # (the first sample takes nvidia-smi a few hundred ms: wait for it, bounded)
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < 3.0:
                time.sleep(0.02)
            self.rows.clear()   # idle samples before the load starts do not count
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.proc = None

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "warm-up + timed steps + end-to-end steps (50 ms period)"}


# ------------------------------------------------------------------ CPU path (oracle port)
def cpu_path(work, steps: int, warmup: int, budget_s: float):
    """The reference's scipy.sparse/SuperLU step (oracle port) on the host cores, from the
    same start state as the b200 arm: `warmup` untimed steps, then `steps` timed ones (cut
    short only if `budget_s` of stepping is exceeded; the count actually timed is returned)."""
    from oracle import tdgl_oracle as orc

    o = work["opts"]
    opts = orc.OracleOptions(solve_time=1e9, dt_init=o["dt_init"], dt_max=o["dt_max"],
                             adaptive=o["adaptive"])
    cf = (lambda t, _c=work["currents"]: _c) if work["currents"] else None
    t0 = time.perf_counter()
    solver = orc.OracleSolver(work["mesh"], opts, work["A"], work["eps"],
                              terminal_info=[orc.TerminalInfo(*t) for t in work["terms"]],
                              current_func=cf)
    factor_s = time.perf_counter() - t0
    psi, mu = start_state(work)
    t, i = 0.0, 0
    for _ in range(warmup):
        dt, psi, mu, _, _ = solver.update(i, t, psi, mu)
        t += dt
        i += 1
    done = 0
    t0 = time.perf_counter()
    while done < steps:
        dt, psi, mu, _, _ = solver.update(i, t, psi, mu)
        t += dt
        i += 1
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    el = time.perf_counter() - t0
    return dict(steps=done, seconds=el, steps_per_s=done / el, factor_seconds=factor_s,
                warmup=warmup)


def threads_used() -> int:
    # scipy's CSR/CSC matvec, SuperLU's triangular solves and the NumPy elementwise code
    # of this path are single-threaded whatever OMP_NUM_THREADS says
    return 1


# ------------------------------------------------------------------ main arms
CPU_FEASIBLE = ("film1m_holes_transport", "film250k_field", "film20_cpu")


def run_reference(args, rank, world):
    if rank != 0:
        return
    # The SuperLU factorisation of the reference's path needs ~10 GB and 25-40 s per million
    # sites (fill grows faster than linearly): a multi-GPU configuration is timed on a bounded
    # SAMPLE of it, the ~1M-site member of the same geometry family (site-steps/s is the unit
    # that carries over); `config` names the configuration, `cpu_baseline.sample` the sample.
    sampled = args.workload not in CPU_FEASIBLE
    name = "film1m_holes_transport" if sampled else args.workload
    work = build_workload(name)
    n = len(work["mesh"].sites)
    res = cpu_path(work, args.steps, args.warmup, budget_s=args.cpu_budget)
    ncores = os.cpu_count()
    sample = (f"{res['steps']} timed steps (of {args.steps}) after {res['warmup']} warm-up steps of"
              f" the {n}-site workload {name}"
              + (f", the 1-GPU member of the geometry family of {args.workload}" if sampled else "")
              + f" (SuperLU factorisation {res['factor_seconds']:.1f} s and mesh build excluded)")
    value = res["steps_per_s"] * n
    # (`config` is the configuration's, identical to the b200 arm's: for a sampled run the big
    # mesh is built only to name its size)
    config = run_config(build_workload(args.workload) if sampled else work)
    line = {
        "impl": "reference", "metric": "tdgl_site_steps_per_sec", "value": value,
        "unit": "site-steps/s", "n_gpus": args.gpus, "steps": res["steps"],
        "warmup": res["warmup"],
        "ms_per_step": 1e3 / res["steps_per_s"], "higher_is_better": True,
        "scaling": "strong" if args.mode == "strong" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "steps_per_sec": res["steps_per_s"],
        "cpu_baseline": {"value": value, "unit": "site-steps/s",
                         "steps_per_sec": res["steps_per_s"], "cores": threads_used(),
                         "host_cores": ncores, "kind": "port", "sample": sample,
                         "sample_workload": name, "sample_sites": n},
        "e2e": {"value": value, "unit": "site-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def strong_record(args, rank, world, local_rank, dist, barrier, torch):
    """BASELINE.json's strong-scaling configuration (configs[4]: ~10M-site film in a field) on
    the job's N GPUs AND on one GPU, measured in the same job with the same code: N-GPU
    ms/step (max over ranks), then rank 0 alone steps the same mesh on its GPU while the other
    ranks wait.  speedup = 1-GPU ms/step / N-GPU ms/step."""
    from tdgl_b200 import SolverOptions, TDGLSolver

    name = args.strong_workload
    K, W = args.steps, args.warmup
    work = build_workload(name, rank, barrier)
    n = len(work["mesh"].sites)
    psi0, mu0 = start_state(work)

    def measure(distributed):
        opts = SolverOptions(solve_time=1e9, save_every=max(K, W, 1), cuda_device=local_rank,
                             use_cuda_graph=not args.no_graph, distributed=distributed,
                             **work["opts"])
        t0 = time.perf_counter()
        solver = TDGLSolver.from_dimensionless(
            work["mesh"], opts, A_applied=work["A"], epsilon=work["eps"],
            terminal_info=work["terms"], terminal_currents=work["currents"])
        setup = time.perf_counter() - t0
        eng = solver.engine
        eng.set_state(psi0, mu0)
        solver.update_mu_boundary(0.0)
        if distributed:
            barrier()
        a = eng.advance(W, 1e300, 0, 0.0) if W > 0 else None
        step, t = (a.step, a.time) if a is not None else (0, 0.0)
        if distributed:
            barrier()
        b = eng.advance(K, 1e300, step, t)
        torch.cuda.synchronize()
        ms = torch.tensor([b.device_ms], dtype=torch.float64, device="cuda")
        if distributed:
            barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        eng.close()
        return float(ms.item()) / K, b.mu_iterations / K, setup

    ms_n, its_n, setup_n = measure(True)
    one = None
    if rank == 0:
        one = measure(False)
    barrier()
    if rank != 0:
        return None
    ms_1, its_1, setup_1 = one
    return {"workload": name, "sites": n, "edges": len(work["mesh"].edge_mesh.edges),
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_n,
            "ms_per_step_1gpu": ms_1, "speedup": ms_1 / ms_n,
            "mu_iterations_per_step": its_n, "mu_iterations_per_step_1gpu": its_1,
            "site_steps_per_sec": n / (ms_n / 1e3), "setup_seconds": {"n_gpus": setup_n, "1gpu": setup_1},
            "how": "same job, same code, same start state; 1-GPU time measured by rank 0 on its own"
                   " GPU while the other ranks wait"}


def run_b200(args, rank, world, local_rank):
    import torch

    from tdgl_b200 import SolverOptions, TDGLSolver
    from tdgl_b200.engine import pinned_empty

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    shard = world > 1 and args.mode != "replicas"
    work = build_workload(args.workload, rank, barrier if world > 1 else None)
    mesh = work["mesh"]
    n, n_edges = len(mesh.sites), len(mesh.edge_mesh.edges)
    K, W = args.steps, args.warmup
    opts = SolverOptions(solve_time=1e9, save_every=max(K, W, 1), cuda_device=local_rank,
                         use_cuda_graph=not args.no_graph, distributed=shard, **work["opts"])
    t0 = time.perf_counter()
    solver = TDGLSolver.from_dimensionless(
        mesh, opts, A_applied=work["A"], epsilon=work["eps"], terminal_info=work["terms"],
        terminal_currents=work["currents"])
    setup_s = time.perf_counter() - t0
    eng = solver.engine
    psi0, mu0 = start_state(work)
    eng.set_state(psi0, mu0)
    solver.update_mu_boundary(0.0)
    info0 = eng.info()

    # ---- device-resident throughput -----------------------------------------------------
    barrier()
    clocks = ClockSampler(local_rank).start()   # (stopped after the end-to-end loop)
    a = eng.advance(W, 1e300, 0, 0.0) if W > 0 else None
    step, t = (a.step, a.time) if a is not None else (0, 0.0)
    launches_before = eng.info()["launches"]
    barrier()
    b = eng.advance(K, 1e300, step, t)
    torch.cuda.synchronize()
    barrier()
    launches = eng.info()["launches"] - launches_before
    ms = torch.tensor([b.device_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    jobs = 1 if shard else world          # a sharded job advances ONE mesh, replicas advance N
    steps_per_s = jobs * K / (ms_total / 1e3)
    iters_per_step = b.mu_iterations / K
    sh = eng.shard_info() if shard else None

    # ---- end to end through the reference's step seam --------------------------------------
    # one GPU: TDGLSolver.update with whole-mesh host arrays (the reference's seam, literally);
    # sharded: every rank moves only the entries it owns through tdgl_update_local (the seam of a
    # job whose state is distributed over the ranks' hosts) — no whole-mesh copies or sums
    state = {"step": b.step, "time": b.time, "dt": b.dt}
    if shard:
        own_s, own_e = eng.local_maps()
        n_loc, e_loc = len(own_s), len(own_e)
        p0, m0 = eng.get_state()
        psi_h = pinned_empty(n_loc, np.complex128)
        mu_h = pinned_empty(n_loc, np.float64)
        psi_h[:] = p0[own_s]
        mu_h[:] = m0[own_s]
        del p0, m0
        out = (pinned_empty(n_loc, np.complex128), pinned_empty(n_loc, np.float64),
               pinned_empty(max(e_loc, 1), np.float64), pinned_empty(max(e_loc, 1), np.float64))
        Ke = K
        api = "tdgl_update_local (owned entries per rank, pinned host arrays)"
    else:
        n_loc, e_loc = n, n_edges
        psi_h = pinned_empty(n, np.complex128)
        mu_h = pinned_empty(n, np.float64)
        out = (pinned_empty(n, np.complex128), pinned_empty(n, np.float64),
               pinned_empty(n_edges, np.float64), pinned_empty(n_edges, np.float64))
        p0, m0 = eng.get_state()
        psi_h[:] = p0
        mu_h[:] = m0
        Ke = K if n <= 2_500_000 else min(K, 40)   # (whole-mesh host copies every step)
        api = "TDGLSolver.update (pinned host arrays)"
    for phase in ("warm", "timed"):
        nsteps = 2 if phase == "warm" else Ke
        barrier()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            if shard:
                solver.update_mu_boundary(state["time"])
                info, _ = eng.update_local(psi_h, mu_h, state["step"], state["time"], out)
                dt_new = info.dt
            else:
                dt_new = solver.update(state, None, state["dt"], psi=psi_h, mu=mu_h, out=out).dt
            # the step's outputs are the next step's host inputs, as in Runner._run_stage
            # (swap the pinned buffers instead of copying host -> host)
            psi_h, mu_h, out = out[0], out[1], (psi_h, mu_h, out[2], out[3])
            state = {"step": state["step"] + 1, "time": state["time"] + dt_new, "dt": dt_new}
        barrier()
        e2e_s = time.perf_counter() - t0
    clocks.stop()
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_steps_per_s = jobs * Ke / float(e2e_t.item())
    # what the link itself can do (pinned 64 MiB copies, best of 5): the floor of e2e
    pcie = {}
    try:
        hb = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        db = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
        for label, src, dst in (("h2d_GBps", hb, db), ("d2h_GBps", db, hb)):
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dst.copy_(src, non_blocking=True)
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            pcie[label] = (64 << 20) / best / 1e6
        del hb, db
    except Exception as exc:
        pcie = {"unavailable": str(exc)}
    # psi (16 B) + mu (8 B) per site in; psi', mu' + J_s, J_n (8 B per edge each) out: whole job
    # (sharded: every site / edge is moved by exactly one rank)
    h2d = 24 * n
    d2h = 24 * n + 16 * n_edges

    # ---- the same measurement deep in the run (vortices nucleating and moving) ---------------
    # the headline numbers start from a state both arms can construct; this record shows that
    # they are not flattered by it: `developed_steps` more steps on the device, then K timed
    developed = None
    if args.developed_steps > 0:
        barrier()
        pre = eng.advance(args.developed_steps, 1e300, state["step"], state["time"])
        barrier()
        dv = eng.advance(K, 1e300, pre.step, pre.time)
        barrier()
        dms = torch.tensor([dv.device_ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(dms, op=dist.ReduceOp.MAX)
        developed = {"pre_roll_steps": int(args.developed_steps + W + K + 2 + Ke),
                     "time_reached": pre.time, "steps": K, "ms_per_step": float(dms.item()) / K,
                     "steps_per_sec": jobs * K / (float(dms.item()) / 1e3),
                     "mu_iterations_per_step": dv.mu_iterations / K, "retries": dv.retries}

    # ---- strong scaling on the 10M-site configuration, in the same job (N > 1) -----------------
    # (rank 0 first times its kernels one by one for the roofline table, the others wait for it
    # in the record's first barrier)
    strong = None
    if rank != 0:
        if shard and args.strong_record:
            eng.close()
            strong_record(args, rank, world, local_rank, dist, barrier, torch)
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- per-kernel roofline (rank 0, kernels timed alone after an L2 flush) ---------------
    peak, peak_src = hbm_peak()
    nnz = info0["nnz"]                      # of this rank's rows
    nl = sh["n_owned"] if shard else n      # rows this rank's kernels process
    levels = info0["amg_levels"]
    kern = {
        "kw_psi_step": (0, 20 * nnz + 52 * nl, 1.0),
        # (complex + real matrix, psi, areas, bterm, mu / mu_prev / mu_pp, b, r, d1, d2)
        "kw_mu_rhs": (1, 28 * nnz + 92 * nl, 1.0),
        # w = A z with r.z and z.w: matrix (8 + 4 B per entry), row pointers 4, z (f32) 4, r 8
        # read, w 8 written per row
        "kw_real<spmv_cg> fine level": (2, 12 * nnz + 24 * nl, iters_per_step),
        # p, s, x, r updated from z (f32), w: 4 + 5 x 8 B read, 4 x 8 B written per row, + 1/diag
        # read and x0 = omega D^-1 r written (4 B each)
        "k_cg_fused": (9, 84 * nl, iters_per_step),
    }
    if levels > 1:
        # the V-cycle's operators are stored in fp32: 4 B value + 4 B column per entry;
        # pre-smoother = residual of x0 (written by k_cg_fused): row pointers 4, r (f64) 8, x0 4
        # read, r1 (f32) 4 written per row;
        # post-smoother: row pointers 4, r 8, 1/diag 4, x 4 read, z (f32) 4 written per row
        kern["kw_real<residual> fine level (pre-smoother)"] = (5, 8 * nnz + 20 * nl, iters_per_step)
        kern["kw_real<jacobi> fine level"] = (6, 8 * nnz + 24 * nl, iters_per_step)
    table = {}
    for name, (which, nbytes, per_step) in kern.items():
        kms = eng.time_kernel(which, 20, flush_l2=True)
        table[name] = {"ms": kms, "bytes": nbytes, "GBps": nbytes / kms / 1e6,
                       "frac": nbytes / kms / 1e6 / peak, "launches_per_step": per_step,
                       "share_of_step": per_step * kms / (ms_total / K)}
    vc_ms = eng.time_kernel(3, 10, flush_l2=True) if levels > 1 else None
    # comparator: cuSPARSE's generic CSR SpMV on the same device arrays
    cusparse = None
    if world == 1:
        try:
            cusparse = {}
            for label, which, nbytes, ours in (
                    ("real_f64", 0, 12 * nnz + 20 * nl, "kw_real<spmv_cg> fine level"),
                    ("complex128", 1, 20 * nnz + 36 * nl, "kw_psi_step")):
                cms = eng.time_cusparse(which, 20, flush_l2=True)
                cusparse[label] = {"ms": cms, "bytes": nbytes, "GBps": nbytes / cms / 1e6,
                                   "frac": nbytes / cms / 1e6 / peak,
                                   "ours": ours, "ours_ms": table[ours]["ms"],
                                   "note": "plain y = A x; ours also fuses the dot products /"
                                           " the psi update into the same pass"}
        except Exception as exc:  # the comparator is optional (library not present)
            cusparse = {"unavailable": str(exc)}
    dom = max(table, key=lambda k: table[k]["share_of_step"])
    # traffic: DRAM bytes need a profiler pass (ncu --set full), which a timed run may not be
    # under; the captures of this command are summarised in profiles/ (r2_full.md)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": table[dom]["GBps"], "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": table[dom]["frac"],
                "traffic": None, "kernels": table, "vcycle_ms": vc_ms,
                "cusparse_spmv": cusparse}

    # ---- the same kernels INSIDE the stepping loop (one GPU) ------------------------------------
    # A second engine created under TDGL_B200_TRACE lets the kernels of the CG iteration write
    # %globaltimer inside the captured CUDA graph (first CTA past griddepcontrol.wait -> last CTA
    # out, tdgl_get_trace): back to back, warm L2, programmatic dependent launch — what a step
    # of `value` consists of, which events around a single launch cannot see.  The timeline
    # costs ~5 % (2.02 -> 2.14 ms/step), so the timed engine above runs without it.
    in_loop = None
    if world == 1 and not args.no_in_loop:
        try:
            os.environ["TDGL_B200_TRACE"] = "2"
            s2 = TDGLSolver.from_dimensionless(
                mesh, opts, A_applied=work["A"], epsilon=work["eps"], terminal_info=work["terms"],
                terminal_currents=work["currents"])
        finally:
            os.environ.pop("TDGL_B200_TRACE", None)
        e2 = s2.engine
        e2.set_state(psi0, mu0)
        s2.update_mu_boundary(0.0)
        tr_info = e2.advance(W + K, 1e300, 0, 0.0)
        tr = e2.get_trace()
        e2.close()
        if tr:
            fine = {"spmv_cg": "kw_real<spmv_cg> fine level", "cg_fused": "k_cg_fused",
                    "residual": "kw_real<residual> fine level (pre-smoother)",
                    "jacobi": "kw_real<jacobi> fine level"}
            rows = {}
            for t in tr:
                if t["rows"] == nl and t["name"] in fine:
                    us = t["t_out"] - t["t_go"]
                    nb = table[fine[t["name"]]]["bytes"]
                    rows[fine[t["name"]]] = {"us": us, "bytes": nb, "GBps": nb / us / 1e3,
                                             "frac": nb / us / 1e3 / peak}
            in_loop = {
                "method": "%globaltimer inside the captured CUDA graph of a second, traced engine"
                          " (TDGL_B200_TRACE, tdgl_get_trace): first CTA past griddepcontrol.wait"
                          " -> last CTA out, last pass through the loop body",
                "iteration_us": max(t["t_out"] for t in tr) - min(t["t_in"] for t in tr),
                "launches_per_iteration": len(tr),
                "ms_per_step_with_trace": tr_info.device_ms / (W + K),
                "kernels": rows,
                "timeline": [[t["name"], t["rows"], round(t["t_in"], 2), round(t["t_go"], 2),
                              round(t["t_out"], 2)] for t in tr]}
    roofline["in_loop"] = in_loop

    if shard and args.strong_record:
        eng.close()
        strong = strong_record(args, rank, world, local_rank, dist, barrier, torch)

    line = {
        "metric": "tdgl_site_steps_per_sec", "value": steps_per_s * n, "unit": "site-steps/s",
        "steps_per_sec": steps_per_s,
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
        "higher_is_better": True, "scaling": "strong" if args.mode == "strong" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": run_config(work),
        "detail": {"nnz_rank0": nnz,
                   "parallelism": ("single GPU" if world == 1 else
                                   f"domain decomposition over {world} GPUs (Z-order ranges,"
                                   " halo exchange + all-reduce kernels on NVLink peer memory)"
                                   if shard else f"{world} replicas"),
                   "shard_rank0": sh,
                   "l2": "operators (>= 230 MB at 1M sites) exceed the 126 MB L2; per-kernel"
                         " roofline timings flush L2 before every launch",
                   "mu_rtol": opts.mu_rtol, "amg_levels": levels},
        "mu_iterations_per_step": iters_per_step, "retries": b.retries,
        "developed": developed,
        "strong": strong,
        "setup_seconds": {"mesh": work["mesh_seconds"], "engine": setup_s},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_steps_per_s * n, "unit": "site-steps/s",
                "steps_per_sec": e2e_steps_per_s, "steps": Ke, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "api": api,
                "per_rank": {"sites": n_loc, "edges": e_loc} if shard else None,
                "link": pcie,
                "link_floor_ms": ((h2d / pcie["h2d_GBps"] + d2h / pcie["d2h_GBps"]) / 1e6
                                  if "h2d_GBps" in pcie else None)},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu:
        res = cpu_path(work, args.cpu_steps, W, budget_s=args.cpu_budget)
        line["cpu_baseline"] = {
            "value": res["steps_per_s"] * n, "unit": "site-steps/s",
            "steps_per_sec": res["steps_per_s"], "cores": threads_used(),
            "host_cores": os.cpu_count(), "kind": "port",
            "sample": (f"{res['steps']} timed steps after {res['warmup']} warm-up steps of the"
                       f" same {n}-site workload from the same start state (SuperLU"
                       f" factorisation {res['factor_seconds']:.1f} s and mesh build excluded)")}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--cpu-budget", type=float, default=240.0,
                    help="seconds of CPU stepping allowed for the cpu baseline / reference arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-in-loop", action="store_true",
                    help="skip the in-loop timeline of the CG iteration's kernels (second engine)")
    ap.add_argument("--developed-steps", type=int, default=2000,
                    help="untimed device steps before the `developed` re-measurement (0: skip)")
    ap.add_argument("--mode", default="weak", choices=["weak", "strong", "replicas"],
                    help="N > 1: domain decomposition of BASELINE.json's config for N GPUs"
                         " (weak, default), of the 1-GPU workload (strong), or replicas")
    ap.add_argument("--no-strong-record", dest="strong_record", action="store_false",
                    help="N > 1: skip the strong-scaling record (10M-site film on N GPUs and on 1)")
    ap.add_argument("--strong-workload", default="film10m_field", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true",
                    help="host-driven launches instead of the device-side-loop CUDA graph"
                         " (profiler runs: every kernel is an ordinary launch)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    if args.workload == "auto":
        args.workload = auto_workload(max(world, args.gpus), args.mode)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        import __graft_entry__ as ge

        ge.build()  # no-op when the in-tree library is current; ranks serialise on a lock
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
