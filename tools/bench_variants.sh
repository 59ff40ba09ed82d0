#!/bin/bash
# Runs bench.py under several engine knobs (env) and prints one summary line per variant.
# usage: tools/bench_variants.sh "NAME1:ENV1=V ENV2=V" "NAME2:..." ...
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  ( env $envs timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu 2>gpurun_out/var_$name.err | tail -1 > gpurun_out/var_$name.json )
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/var_{name}.json"))
    k = d["roofline"]["kernels"]
    ks = " ".join(f"{n.split(' fine')[0].replace('kw_real','')}={v['ms']*1e3:.1f}us/{v['frac']:.2f}" for n, v in k.items())
    dv = d.get("developed") or {}
    print(f"{name}: ms/step={d['ms_per_step']:.3f} its={d['mu_iterations_per_step']:.2f} dev_ms={dv.get('ms_per_step', 0):.3f} dev_its={dv.get('mu_iterations_per_step', 0):.2f} e2e={d['e2e']['steps_per_sec']:.1f}/s vcycle={d['roofline']['vcycle_ms']*1e3:.1f}us levels={d['detail']['amg_levels']} setup={d['setup_seconds']['engine']:.1f}s | {ks}")
except Exception as e:
    print(name, "FAILED", e, open(f"gpurun_out/var_{name}.err").read()[-1500:])
PY
done
