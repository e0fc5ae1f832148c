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
        if k.startswith("void "):
            k = k[5:]
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


def _f(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return 0.0


def source(src, dst, top=10):
    """Hot source lines per kernel (warp-stall samples, executed warp instructions, dominant stall reason) from an
    --import-source capture; SASS rows of the source page are aggregated per CUDA-C line."""
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    names = []
    for r in rows[2:]:
        k = r[rows[0].index("Kernel Name")].split("(")[0].replace("void ", "")
        if k not in names and k.startswith("k_"):
            names.append(k)
    out = [f"# ncu source page: `{src}` (hot CUDA-C lines; share of warp-stall samples / of executed warp instructions)", ""]
    for k in names:
        pat = k.split("<")[0]
        txt = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{pat}"],
                             capture_output=True, text=True).stdout
        agg, hdr, fname = {}, None, "?"
        for r in csv.reader(txt.splitlines()):
            if not r:
                continue
            if r[0] == "File Name":
                fname = r[1].split("/")[-1]; continue
            if r[0] == "Line No":
                hdr = r; continue
            if hdr is None or len(r) < len(hdr):
                continue
            try:
                ln = int(r[0])
            except ValueError:
                continue
            if "# Samples" not in hdr or "Instructions Executed" not in hdr:
                continue
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            a = agg.setdefault((fname, ln), [r[1].strip(), 0.0, 0.0, collections.Counter()])
            a[1] += _f(r[i_s]); a[2] += _f(r[i_i])
            for i in st:
                a[3][hdr[i]] += _f(r[i])
        tot = sum(v[1] for v in agg.values()) or 1.0
        toti = sum(v[2] for v in agg.values()) or 1.0
        out += [f"## `{k}`", "", f"{int(tot)} samples, {int(toti)} warp instructions", "",
                "| file:line | samples % | instr % | top stall | source |", "|---|---|---|---|---|"]
        for (fn, ln), (text, smp, ins, stl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
            ts = stl.most_common(1)[0][0] if stl else ""
            out.append(f"| {fn}:{ln} | {100 * smp / tot:.1f} | {100 * ins / toti:.1f} | {ts} | `{text[:100].replace('|', '/')}` |")
        out.append("")
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    {"launches": launches, "full": full, "source": source}[mode](src, dst)
    print("wrote", dst)
