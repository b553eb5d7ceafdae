"""Summarise ncu artefacts from gpurun_out/ into tracked files under profiles/.
  python tools/summarize_ncu.py launches <launches.csv> <out.md>
  python tools/summarize_ncu.py kernel <file.ncu-rep> <out.md> [title]
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "gpc__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as fh:
        fh.write(f"# ncu launch list ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
                 "(cold-cache, serialised: compare SHARES)\n\n")
        fh.write("| kernel | launches | total ms | avg us | share | last grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / a[0] / 1e3:.1f} | {a[1] / tot:.4f} | {a[2]} | {a[3]} |\n")
    print(open(dst).read())


def kernel(src, dst, title):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as fh:
        fh.write(f"# {title}\n\nsource: `{src}` (`ncu --set full --clock-control none --import-source on`)\n\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            fh.write(f"## `{d.get('Kernel Name', '')[:100]}` grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n")
            fh.write("| metric | unit | value |\n|---|---|---:|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    fh.write(f"| {k} | {units[i]} | {vals[i]} |\n")
            fh.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else sys.argv[2])
