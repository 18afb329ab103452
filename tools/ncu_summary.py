#!/usr/bin/env python
"""Key metrics of the first kernel in an .ncu-rep as a markdown table (read here, on the CPU box).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "avg active threads / instruction"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads"),
    ("smsp__inst_executed_op_local_st.sum", "local stores"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, r = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    print("kernel: `%s`\n" % r[col["Kernel Name"]])
    print("| metric | value |\n|---|---|")
    for key, label in WANT:
        if key in col:
            print("| %s (`%s`) | %s %s |" % (label, key, r[col[key]], units[col[key]]))
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warp_latency_issue_stalled_") or h.startswith("smsp__average_warps_issue_stalled_"):
            if h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                name = h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")
                try:
                    stalls.append((float(r[i]), name))
                except ValueError:
                    pass
    stalls.sort(reverse=True)
    print("| stalls per issue (top) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]))


if __name__ == "__main__":
    main()
