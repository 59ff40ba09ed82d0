#!/usr/bin/env python
"""Writes profiles/r2_sass.txt: for the hot kernels of the step, the SASS lines that show the
TMA bulk copies (UBLKCP), the mbarrier waits (SYNCS.PHASECHK), the programmatic-dependent-launch
instructions (ACQBULK / griddepcontrol) and the fp64 FMA pipeline (DFMA), taken with
`cuobjdump -sass` from the built library.  Run after __graft_entry__.build()."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "py-tdgl_b200", "libtdgl_b200.so")
WANT = [("kw_real<spmv_cg, single GPU, 1 lane/row, double matrix, float z>", r"kw_realILi6ELb0ELi1ENS_9RealTypesIdfddd"),
        ("kw_real<residual (pre-smoother), single GPU, 1 lane/row, float matrix, double r>", r"kw_realILi1ELb0ELi1ENS_9RealTypesIffdff"),
        ("kw_real<jacobi, single GPU, 1 lane/row, float matrix, double r, float z>", r"kw_realILi3ELb0ELi1ENS_9RealTypesIffdff"),
        ("kw_real<restriction, single GPU, 4 lanes/row, float>", r"kw_realILi4ELb0ELi4ENS_9RealTypesIfffff"),
        ("kw_psi_step<single GPU>", r"kw_psi_stepILb0E"),
        ("kw_mu_rhs<single GPU>", r"kw_mu_rhsILb0E"),
        ("kw_real<spmv_cg, sharded>", r"kw_realILi6ELb1ELi1ENS_9RealTypesIdfddd")]
KEYS = ("UBLKCP", "SYNCS", "ACQBULK", "DEPBAR", "ST.E.64.STRONG.SYS", "LD.E.128.STRONG.SYS",
        "LD.E.64.STRONG.SYS")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    out = ["# SASS excerpts of libtdgl_b200.so (sm_100a), `cuobjdump -sass`", ""]
    total = {k: sass.count(k) for k in ("UBLKCP", "SYNCS.PHASECHK", "DFMA")}
    out.append(f"whole library: {total['UBLKCP']} UBLKCP (cp.async.bulk global->shared), "
               f"{total['SYNCS.PHASECHK']} SYNCS.PHASECHK (mbarrier try_wait), {total['DFMA']} DFMA")
    out.append("")
    for title, pat in WANT:
        for f in funcs:
            name = f.split("\n", 1)[0]
            if re.search(pat, name):
                lines = f.split("\n")
                n_dfma = sum("DFMA" in l for l in lines)
                n_ldg = sum(re.search(r"\bLDG", l) is not None for l in lines)
                n_lds = sum(re.search(r"\bLDS", l) is not None for l in lines)
                out.append(f"## {title}")
                out.append(f"   {name.strip()}")
                out.append(f"   {len(lines)} lines; DFMA {n_dfma}, LDG {n_ldg}, LDS {n_lds}")
                for l in lines:
                    if any(k in l for k in KEYS):
                        out.append("   " + re.sub(r"\s+", " ", l.strip())[:150])
                out.append("")
                break
        else:
            out.append(f"## {title}: NOT FOUND")
    path = os.path.join(ROOT, "profiles", "r2_sass.txt")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("wrote", path, len(out), "lines")


if __name__ == "__main__":
    sys.exit(main())
