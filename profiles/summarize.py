"""Turns ncu outputs brought back under gpurun_out/ into the small text/JSON summaries kept in
profiles/.  Usage:
  python profiles/summarize.py launches gpurun_out/launches_X.csv profiles/launches_X_summary.txt "<command line>"
  python profiles/summarize.py full gpurun_out/prof_X.ncu-rep profiles/ncu_full_X.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
]


def launches(src, dst, cmd):
    rows = list(csv.reader(open(src)))
    h = next(k for k, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mi:
            continue
        name = r[ki].split("(")[0]
        v = float(r[mi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none: {cmd}\n")
        f.write("(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)\n")
        f.write(f"{'kernel':62s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{n:62s} {a[0]:8d} {a[1]:12.1f} {a[1] / a[0]:10.2f} {100 * a[1] / tot:6.2f}%\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"ncu --set full --clock-control none --import-source on  ({src}); selected raw metrics per launch\n")
        for r in rows[2:]:
            f.write("\n" + r[idx["Kernel Name"]] + "\n")
            for k in KEEP:
                if k in idx:
                    f.write(f"    {k:75s} {r[idx[k]]:>18s} {units[idx[k]]}\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
