#!/usr/bin/env python
"""Turns ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_rNN.csv profiles/launches_rNN.md
      (from: ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file ... python bench.py ...)
  python tools/summarize_ncu.py full gpurun_out/prof_rNN.ncu-rep profiles/ncu_full_rNN.md
      (from: ncu --set full --clock-control none --import-source on -k regex:... -o ... python bench.py ...)
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        a = agg.setdefault(k, [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(v)
    ours = {k: v for k, v in agg.items() if k.startswith("rvio::")}
    tot = sum(v[1] for v in ours.values())
    out = ["# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)", "",
           f"source: `{src}`; {sum(v[0] for v in ours.values())} launches of this library's kernels, {tot:.1f} us in total", "",
           "| kernel | launches | avg us | median us | share of kernel time |", "|---|---|---|---|---|"]
    for k, (c, t, vs) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        vs = sorted(vs)
        out.append(f"| `{k}` | {c} | {t / c:.2f} | {vs[len(vs) // 2]:.2f} | {100 * t / tot:.1f}% |")
    other = {k: v for k, v in agg.items() if not k.startswith("rvio::")}
    if other:
        out += ["", "Other kernels in the capture (bench harness: torch L2-flush fill etc.): " +
                ", ".join(f"`{k[:60]}` x{v[0]}" for k, v in other.items())]
    open(dst, "w").write("\n".join(out) + "\n")


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_static", "static smem/block"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: mio throttle"),
]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full capture: `{src}`", "", "(per launch; units as reported by ncu)", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        out += [f"## `{name}`", "", "| metric | value |", "|---|---|"]
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                out.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
        out.append("")
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    (launches if mode == "launches" else full)(src, dst)
    print("wrote", dst)
