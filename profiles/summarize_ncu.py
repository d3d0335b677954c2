#!/usr/bin/env python
"""Condense ncu output into the text summaries kept under profiles/.

    python profiles/summarize_ncu.py full  gpurun_out/prof_X.ncu-rep  > profiles/rNN_X_full.txt
    python profiles/summarize_ncu.py list  gpurun_out/launches.csv    > profiles/rNN_X_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none, source: {path}")
    for k, r in enumerate(rows[2:]):
        print(f"\n## launch {k}")
        for name in KEEP:
            if name in hdr:
                i = hdr.index(name)
                print(f"{name:82s} {r[i]} {units[i]}")


def launch_list(path):
    text = open(path).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    agg = OrderedDict()
    seq = []
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        seq.append((name, v))
    # the step region = everything after the last launch that is not ours
    # (bench.py generates its synthetic input with torch first, then only
    # launches libnfm kernels: warm-up steps + timed steps)
    last_setup = max([i for i, (n, _) in enumerate(seq) if "nfm::" not in n], default=-1)
    region = seq[last_setup + 1:]
    total = sum(a[1] for a in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none, source: {path}")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"# {sum(a[0] for a in agg.values())} launches, {total:.1f} us total\n")
    print(f"{'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}  kernel")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{cnt:8d} {us:12.1f} {us / cnt:10.2f} {us / total:7.1%}  {name[:150]}")
    rt = sum(v for _, v in region)
    print(f"\n# step region (warm-up + timed steps; after the last torch set-up launch): {len(region)} launches, {rt:.1f} us")
    ragg = OrderedDict()
    for n, v in region:
        a = ragg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    for name, (cnt, us) in ragg.items():
        print(f"{cnt:8d} {us:12.1f} {us / cnt:10.2f} {us / max(rt, 1e-9):7.1%}  {name[:150]}")


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2])
