#!/usr/bin/env python
"""Summarise an `ncu --set full` capture (.ncu-rep, or the CSV of `ncu -i X --page raw --csv`) into
the markdown kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/rNN_<what>.md
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration (under ncu, cold, serialised)"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit by registers (blocks/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("lts__t_bytes.sum", "L2 bytes (all)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 sectors, global loads"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 requests, global loads"),
    ("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "L1 requests, local loads (stack)"),
    ("l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "L1 requests, local stores (stack)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (memory)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected (issue-limited)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
]


def load(path):
    if path.endswith(".ncu-rep"):
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True, check=True).stdout
        rows = list(csv.reader(txt.splitlines()))
    else:
        rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    return [(dict(zip(hdr, r)), dict(zip(hdr, units))) for r in rows[2:]]


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    print(f"# {title}\n")
    print(f"Source: `{path}` (`ncu --set full --clock-control none --import-source on`, one B200, "
          "under gpurun).  Durations under ncu are cold-cache and serialised; the bench numbers are "
          "taken without a profiler.\n")
    for vals, units in load(path):
        print(f"## `{vals.get('Kernel Name', '?')}`  (launch id {vals.get('ID', '?')})\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k, label in KEYS:
            if k in vals and vals[k] != "":
                print(f"| {label} (`{k}`) | {vals[k]} | {units.get(k, '')} |")
        try:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            rd = float(vals["dram__bytes_read.sum"].replace(",", "")) * \
                scale[units.get("dram__bytes_read.sum", "byte")]
            wr = float(vals["dram__bytes_write.sum"].replace(",", "")) * \
                scale[units.get("dram__bytes_write.sum", "byte")]
            print(f"\nDRAM traffic (read + write) per launch: {(rd + wr) / 1e6:.3f} MB\n")
        except Exception:
            print()


if __name__ == "__main__":
    main()
