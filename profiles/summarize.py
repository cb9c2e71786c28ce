#!/usr/bin/env python
"""Turn ncu outputs brought back from the GPU box into the small text summaries kept in profiles/.

    python profiles/summarize.py full    gpurun_out/<name>.ncu-rep  > profiles/<name>_full.txt
    python profiles/summarize.py launches gpurun_out/<name>.csv     > profiles/<name>_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {path} (per launch; cold-cache, serialised replays)")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"\nkernel: {name}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                print(f"  {k:72s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        for k in hdr:
            if "stall" in k and k.endswith("_per_issue_active.ratio") is False and "issue_stalled" in k and "pct" in k:
                print(f"  {k:72s} {r[hdr.index(k)]}")


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) >= 15 and r[0].isdigit()]
    agg = OrderedDict()
    total = 0.0
    print(f"# launch list from {path}: ncu --metrics gpu__time_duration.sum --clock-control none (ns, cold-cache, serialised)")
    for r in rows:
        try:
            ns = float(r[14].replace(",", ""))
        except ValueError:
            continue
        name = r[4].split("(")[0][-90:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e3:12.1f} us  {100 * ns / total:5.1f} %  x{n:<4d} {name}")
    print(f"{total / 1e3:12.1f} us  total over {len(rows)} launches")


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2])
