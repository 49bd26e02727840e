#!/usr/bin/env python3
"""Turn gpurun_out/ ncu artefacts into the text summaries kept under profiles/ (run in the build container)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum"]
STALLS = "smsp__average_warps_issue_stalled_"


def launches(path: str, last: int) -> str:
    rows = [r for r in csv.reader(open(path, encoding="utf-8")) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    rows = rows[-last:]
    total = sum(float(r[vi]) for r in rows)
    out = [f"# per-launch gpu__time_duration.sum (ns), last {last} launches = one pass; cold-cache, serialised: compare SHARES"]
    for r in rows:
        out.append(f"{r[ki][:90]:90s} {float(r[vi]):10.0f} {100 * float(r[vi]) / total:5.1f}%")
    out.append(f"{'total':90s} {total:10.0f}")
    return "\n".join(out)


def report(path: str, kernel: str = "") -> str:
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    vals = [r for r in rows[2:] if kernel in r[hdr.index("Kernel Name")]][-1]   # the last captured launch of that kernel
    out = [f"# ncu --set full --clock-control none, one launch: {vals[hdr.index('Kernel Name')][:100]}"]
    for h, u, v in zip(hdr, units, vals):
        if h in WANT or (h.startswith(STALLS) and h.endswith("per_issue_active.ratio")):
            out.append(f"{h:90s} {v} {u}")
    return "\n".join(out)


if __name__ == "__main__":
    mode, src = sys.argv[1], sys.argv[2]
    print(launches(src, int(sys.argv[3])) if mode == "launches" else report(src, sys.argv[3] if len(sys.argv) > 3 else ""))
