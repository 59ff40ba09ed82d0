"""Turn the ncu outputs of a gpurun call into the committed summaries under profiles/.

    python tools/ncu_summary.py <tag> [launches.csv] [prof.ncu-rep]

  profiles/<tag>_launches.csv      the launch list as ncu wrote it (gpu__time_duration.sum)
  profiles/<tag>_launches.md       per-kernel count / total / mean / share of the step
  profiles/<tag>_full.md           selected `--set full` metrics per captured launch
"""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

OPS = {"kw_real<0>": "kw_real<spmv_dot>", "kw_real<1>": "kw_real<residual>",
       "kw_real<2>": "kw_real<presmooth>", "kw_real<3>": "kw_real<jacobi>",
       "kw_real<4>": "kw_real<plain> (restriction)", "kw_real<5>": "kw_real<plain_add> (prolongation)",
       "kw_real<6>": "kw_real<spmv_cg>"}


def short(name):
    name = name.replace("void ", "").replace("tdgl::", "").strip()
    m = re.match(r"kw_real<\(?(?:int\))?(\d), \(?(?:bool\))?(\d)(?:, \(?(?:int\))?(\d))?(?:, RealTypes<([^>]*)>)?", name)
    if m:
        types = (m.group(4) or "").replace(" ", "")
        tag = {"double,double,double,double,double": "f64", "float,float,float,float,float": "f32",
               "float,float,double,float,float": "f32 matrix, f64 rhs",
               "double,float,double,double,double": "f64 matrix, f32 x"}.get(types, types)
        return (OPS[f"kw_real<{m.group(1)}>"] + (" [sharded]" if m.group(2) == "1" else "")
                + (" [4 lanes/row]" if m.group(3) == "4" else "") + (f" [{tag}]" if tag else ""))
    name = re.sub(r"\([^()]*\)\s*$", "", name)           # the argument list
    name = re.sub(r"\((?:int|bool)\)", "", name)
    return OPS.get(name, name)


if os.path.exists(launches):
    shutil.copy(launches, os.path.join(out, f"{tag}_launches.csv"))
    with open(launches) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}[row["Metric Unit"]]
        key = (short(row["Kernel Name"]), row["Grid Size"], row["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    bench = {k: v for k, v in agg.items() if k[0] != "k_flush_l2"}
    tot_b = sum(v[1] for v in bench.values())
    with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, --clock-control none\n\n"
                "Per-launch times under ncu are cold-cache and serialised: read the SHARES.\n"
                "`k_flush_l2` (the L2 flush of the per-kernel timer in bench.py) is excluded from the shares.\n"
                "Rows are (kernel, grid, block): the same CSR window kernel appears once per AMG level.\n\n"
                "| kernel | grid | block | launches | total us | mean us | share |\n|---|---|---|---|---|---|---|\n")
        for (k, g, b), (c, t) in sorted(bench.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {g} | {b} | {c} | {t:.1f} | {t / c:.2f} | {t / tot_b:.3f} |\n")
        by_kernel = collections.OrderedDict()
        for (k, g, b), (c, t) in bench.items():
            a = by_kernel.setdefault(k, [0, 0.0])
            a[0] += c
            a[1] += t
        f.write("\n## by kernel (all levels)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (c, t) in sorted(by_kernel.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {c} | {t:.1f} | {t / tot_b:.3f} |\n")
        f.write(f"\ntotal {tot_b / 1e3:.2f} ms over {sum(v[0] for v in bench.values())} launches\n")
    print("wrote", os.path.join(out, f"{tag}_launches.md"))

if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(out, f"{tag}_full.md"), "w") as f:
        f.write(f"# ncu --set full ({tag}), selected metrics per captured launch\n\n"
                "traffic = dram__bytes_read.sum + dram__bytes_write.sum (per launch)\n\n")
        f.write("| " + " | ".join(f"{hdr[i]} [{units[i]}]" for i in idx) + " | traffic MB | GB/s |\n")
        f.write("|" + "---|" * (len(idx) + 2) + "\n")
        for r in rows[2:]:
            vals = [short(r[i]) if hdr[i] == "Kernel Name" else r[i] for i in idx]
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
                wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
                us = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
                sc = {"Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}
                rd *= sc[units[hdr.index("dram__bytes_read.sum")]]
                wr *= sc[units[hdr.index("dram__bytes_write.sum")]]
                us *= {"us": 1.0, "ns": 1e-3, "ms": 1e3}[units[hdr.index("gpu__time_duration.sum")]]
                extra = [f"{rd + wr:.2f}", f"{(rd + wr) / us * 1e3:.0f}"]
            except Exception:
                extra = ["", ""]
            f.write("| " + " | ".join(vals + extra) + " |\n")
    print("wrote", os.path.join(out, f"{tag}_full.md"))
