#!/usr/bin/env python
"""Per-kernel SASS summary of librvio_b200.so (runs without a GPU): instruction counts of the mnemonics that matter for
the Blackwell-evidence table of /opt/skills/guides/B200_PROFILING.md (UTMALDG = TMA, DMMA/HMMA/UTC*MMA = tensor cores,
LDGSTS = cp.async, SYNCS = mbarrier, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents of programmatic dependent
launch, DFMA = FP64 SIMT ...).

    python tools/sass_summary.py > profiles/sass_r02.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "r-vio_b200", "librvio_b200.so")
KEYS = ["DMMA", "HMMA", "UTC", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "LDTM", "ACQBULK", "PREEXIT", "DFMA", "DMUL", "DADD", "FFMA", "IMAD", "LDG", "STG", "LDS", "STS",
        "BAR", "SHFL", "REDUX", "ATOM", "MUFU", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o).replace("rvio::", "") for o in out]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); funcs[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            funcs[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    funcs[cur][k] += 1
    names = demangle(list(funcs))
    print("# SASS summary of `r-vio_b200/librvio_b200.so` (sm_100a), `cuobjdump -sass`, instruction counts per kernel\n")
    print("Tensor-core / TMA evidence: `DMMA` = FP64 tensor-core MMA (`mma.sync.m8n8k4.f64`), `UTMALDG` = TMA tensor load "
          "(`cp.async.bulk.tensor`), `SYNCS` = mbarrier operations, `LDGSTS` = `cp.async`, `ACQBULK` / `PREEXIT` = `griddepcontrol.wait` / "
          "`griddepcontrol.launch_dependents` (programmatic dependent launch, DESIGN.md section 5).\n")
    cols = ["DMMA", "UTMALDG", "SYNCS", "LDGSTS", "ACQBULK", "PREEXIT", "DFMA", "DMUL", "FFMA", "LDG", "STG", "LDS", "STS", "BAR", "SHFL", "REDUX", "ATOM", "LDL", "STL"]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    tot = collections.Counter()
    for (f, c), nm in sorted(zip(funcs.items(), names), key=lambda t: t[1]):
        print(f"| `{nm}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
        tot.update(c)
    print(f"| **all {len(funcs)} kernels** | {tot['_total']} | " + " | ".join(str(tot[k]) if tot[k] else "" for k in cols) + " |")
    print(f"\n`HMMA` {tot['HMMA']}, `UTC*MMA` {tot['UTC']}, `LDTM` {tot['LDTM']}: the path has no FP16/BF16/TF32 GEMM -- its dense contractions are FP64 "
          "(accept / reject decisions and the filter state must match the reference to 1e-9), and tcgen05 has no FP64 kind; see DESIGN.md.")


if __name__ == "__main__":
    main()
