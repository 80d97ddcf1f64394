"""Summary of the ncu CSV pages (`ncu -i x.ncu-rep --page raw|source --csv`, exported on the GPU box by scripts/gpu_final*.sh
because the reports exceed the transfer limit): key metrics, instruction share / stall-sample share of the SASS regions, opcode
histogram.  usage: python profiles/ncu_csv_summary.py <prefix> [work items for per-item instruction counts] > profiles/rNN_x.txt"""
import csv
import sys

pre = sys.argv[1]
ncells = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(pre + ".raw.csv")))
h, u, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "launch__grid_size", "launch__block_size"]
for a, b, c in zip(h, u, v):
    if a in want or a.startswith("smsp__average_warps_issue_stalled") and float(c or 0) > 0.3:
        print(f"{a:90s} {b:12s} {c}")
src = list(csv.reader(open(pre + ".src.csv")))[2:]
tot = sum(int(r[5]) for r in src)
ts = sum(int(r[2]) for r in src)
print("instructions", tot, "per cell", tot / ncells, "samples", ts, "sass lines", len(src))
seg = []
cur = None
for i, r in enumerate(src):
    c = int(r[5])
    if cur and abs(c - cur[2]) <= 0.02 * max(c, cur[2]):
        cur[1] = i; cur[3] += c; cur[4] += int(r[2])
    else:
        cur = [i, i, c, c, int(r[2])]; seg.append(cur)
for s in seg:
    if s[3] > 0.02 * tot or s[4] > 0.02 * ts:
        print(f"sass {s[0]:5d}-{s[1]:5d} n={s[1]-s[0]+1:4d} exec/cell={s[2]/ncells:9.2f} inst share={s[3]/tot:.3f} sample share={s[4]/ts:.3f}")
top = sorted(range(len(src)), key=lambda i: -int(src[i][2]))[:14]
for i in sorted(top):
    print(i, src[i][1].strip()[:80], src[i][5], src[i][2])

import collections
op = collections.Counter()
for r in src:
    toks = r[1].split()
    if not toks:
        continue
    o = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
    op[o] += int(r[5])
print("opcode histogram (warp-level executed):", ", ".join(f"{k} {v}" for k, v in op.most_common(24)))
