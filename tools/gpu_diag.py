"""One-shot GPU diagnostics: per-kernel timings and achieved bandwidth on a synthetic film.
Usage: python tools/gpu_diag.py [size_xi] [steps]   (writes gpurun_out/diag_<N>.json)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tdgl_b200.engine import DeviceEngine  # noqa: E402
from tdgl_b200.synthetic import film_problem  # noqa: E402

size = float(sys.argv[1]) if len(sys.argv) > 1 else 400.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
use_graph = int(os.environ.get("TDGL_GRAPH", "1"))
t0 = time.time()
mesh, A, eps, _ = film_problem(size, size, 0.43, b=0.1)
n = len(mesh.sites)
t1 = time.time()
eng = DeviceEngine(mesh, use_graph=use_graph, running_capacity=max(steps, 16))
t2 = time.time()
info = eng.info()
print(f"N={n} mesh {t1-t0:.1f}s create {t2-t1:.1f}s info={info} msg={eng.message!r}", flush=True)
eng.set_link_exponents(A)
eng.set_epsilon(eps)
eng.set_stepper(dt_init=1e-4, dt_max=1e-1)
out = dict(N=n, info=info, mesh_s=t1 - t0, create_s=t2 - t1)
# warm-up steps (also gives the mu solver a meaningful rhs)
a = eng.advance(20, 1e300, 0, 0.0)
print("warmup", a, flush=True)
ts = time.time()
b = eng.advance(steps, 1e300, a.step, a.time)
te = time.time()
print("timed", b, f"{steps/(te-ts):.1f} steps/s wall", flush=True)
out.update(steps_per_s=steps / (te - ts), mu_iters_per_step=b.mu_iterations / steps,
           retries=b.retries, dt_last=b.dt)
nnz = info["nnz"]
alg = {0: 20 * nnz + 52 * n, 1: 28 * nnz + 60 * n, 2: 12 * nnz + 20 * n, 5: 12 * nnz + 36 * n,
       6: 12 * nnz + 44 * n}
names = {0: "psi_step", 1: "mu_rhs", 2: "mu_spmv_dot", 3: "vcycle", 4: "mu_solve_cold",
         5: "presmooth0", 6: "jacobi0", 7: "restrict0", 8: "prolong0"}
for flush in (True, False):
    for k in range(9):
        ms = eng.time_kernel(k, 20 if k != 4 else 3, flush_l2=flush)
        rec = dict(ms=ms)
        if k in alg:
            rec["GBps"] = alg[k] / ms / 1e6
            rec["frac_of_6535"] = rec["GBps"] / 6535.1
        out[names[k] + ("_flush" if flush else "_b2b")] = rec
        print(names[k], "flush" if flush else "back-to-back", rec, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"diag_{n}_g{use_graph}.json"), "w") as f:
    json.dump(out, f, indent=1)
