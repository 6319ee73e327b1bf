#!/usr/bin/env python
"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).
  python tools/summarize_profiles.py launches <launches.csv> <out.md> [title]
  python tools/summarize_profiles.py rep <file.ncu-rep> <out.md> [title]
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max"]


def launches(path, out, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (row["Kernel Name"].split("(")[0][:80], row["Grid Size"], row["Block Size"])
        agg.setdefault(key, []).append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\nncu `--metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: "
                f"compare SHARES, not absolutes).  Source: `{path}`.\n\n")
        f.write("| kernel | grid | block | launches | avg us | total us | share |\n|---|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[0]}` | {k[1]} | {k[2]} | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / 1e3:.1f} | "
                    f"{100 * sum(v) / tot:.1f}% |\n")
        f.write(f"\nTotal device time of the listed launches: {tot / 1e3:.1f} us\n")


def rep(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on`; source `{path}`.\n\n")
        for row in r[2:]:
            kn = row[hdr.index("Kernel Name")]
            f.write(f"## `{kn[:100]}` grid {row[hdr.index('Grid Size')]} block {row[hdr.index('Block Size')]}\n\n")
            f.write("| metric | value | unit |\n|---|---|---|\n")
            for i, h in enumerate(hdr):
                if h in KEEP:
                    f.write(f"| {h} | {row[i]} | {units[i]} |\n")
            f.write("\n")


if __name__ == "__main__":
    mode, path, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else path
    (launches if mode == "launches" else rep)(path, out, title)
