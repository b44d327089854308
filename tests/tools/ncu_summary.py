#!/usr/bin/env python
"""Turns an .ncu-rep (one kernel, `ncu --set full`) into the short markdown kept under profiles/.
usage: python tests/tools/ncu_summary.py gpurun_out/sc_r1.ncu-rep "title" > profiles/xyz.md"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid size"), ("launch__block_size", "block size"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"), ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n")
    print(f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`, one launch, warm).  Times taken under ncu are\n"
          "serialised and cold-cache; bench.py's CUDA-event numbers are the ones quoted as results.\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## `{d.get('Kernel Name', '?')}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for key, label in WANT:
            if key in d and d[key] != "":
                print(f"| {label} (`{key}`) | {d[key]} | {units[hdr.index(key)]} |")
        print()


if __name__ == "__main__":
    main()
