"""Summarise an .ncu-rep (one or more kernels captured with --set full) as plain text for profiles/.

usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# summary of {rep} (ncu --set full --clock-control none; per-launch, cold cache, serialised)")
    for r in rows[2:]:
        print()
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k} [{units[i]}] = {r[i][:90]}")
        print("warp stalls per issue-active cycle (> 0.05):")
        pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
        for i, h in enumerate(hdr):
            if h.startswith(pre) and h.endswith(suf):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.05:
                    print(f"   {h[len(pre):-len(suf)]} = {v:.2f}")


if __name__ == "__main__":
    main()
