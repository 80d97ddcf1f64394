#!/usr/bin/env python
"""Summarise an .ncu-rep (run where ncu is installed): key raw metrics per launch + opcode histogram.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio']
for vals in rows[2:]:
    out.write("=" * 100 + "\n")
    for h, u, v in zip(hdr, units, vals):
        if h in keys:
            out.write(f"{h} [{u}] = {v}\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name",')
for blk in blocks[1:]:
    r = list(csv.reader(io.StringIO(blk)))
    name = r[0][0]
    h = r[1]
    ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    op = collections.Counter(); sm = collections.Counter()
    for x in r[2:]:
        if len(x) <= max(ie, isamp): continue
        toks = x[ia].split()
        o = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
        op[o] += int(x[ie]); sm[o] += int(x[isamp])
    tot, ts = sum(op.values()), max(1, sum(sm.values()))
    out.write("-" * 100 + f"\nopcode histogram: {name}\n  total warp-instructions {tot}\n")
    for o, c in op.most_common(16):
        out.write(f"  {o:8s} {c:12d} {100*c/tot:5.1f}%   stall-samples {100*sm[o]/ts:5.1f}%\n")
