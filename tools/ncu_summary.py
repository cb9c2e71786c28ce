"""Summarise an ncu report (.ncu-rep) here on the CPU box: key throughput metrics and warp stall reasons per launch.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [extra-metric-substring ...]
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = dict(zip(h, v))
        unit = dict(zip(h, u))
        print("kernel:", d.get("Kernel Name", "")[:100], "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]} {unit[k]}")
        for k in h:
            if any(e in k for e in extra):
                print(f"  {k:75s} {d[k]} {unit[k]}")
        stalls = [(float(d[k]), k) for k in h if "issue_stalled" in k and k.endswith("per_warp_active.pct") and d[k]]
        for val, k in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {k.split('issue_stalled_')[1].split('_per_warp')[0]:40s} {val:.1f} % of warp-active cycles")


if __name__ == "__main__":
    main()
