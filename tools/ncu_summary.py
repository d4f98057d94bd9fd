#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into one line per launch."""
import csv, subprocess, sys
rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = rows[0]
def col(name):
    for i, c in enumerate(h):
        if c == name: return i
    return None
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__inst_executed.sum", "inst"),
        ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("lts__t_sector_hit_rate.pct", "L2hit%"), ("lts__t_bytes.sum", "L2bytes"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"), ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"), ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br"), ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"), ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu%"),
        ("smsp__inst_executed.sum", "warp_inst"),
        ("local_load_bytes", "lld"), ("smsp__inst_executed_op_local_ld.sum", "ld.local"), ("smsp__inst_executed_op_local_st.sum", "st.local")]
idx = [(col(n), s) for n, s in want]
units = rows[1]
for r in rows[2:]:
    out = []
    for i, s in idx:
        if i is None: continue
        v = r[i]
        if s == "kernel": v = v.split("(")[0][:28]
        else:
            u = units[i]
            out_u = {"msecond": "ms", "usecond": "us", "Mbyte": "MB", "Gbyte": "GB", "Kbyte": "KB", "byte": "B"}.get(u, "")
            v = v + out_u
        out.append(f"{s}={v}")
    print("  ".join(out))
