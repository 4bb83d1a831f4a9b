"""Summarises an .ncu-rep: key raw metrics + executed-instruction histogram by opcode and by source line.
usage: python profiles/ncu_summary.py file.ncu-rep [top_lines]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__average_warps_issue_stalled', 'sm__cycles_elapsed.avg.per_second',
        'sm__cycles_active.avg', 'smsp__sass_inst_executed_op_local', 'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate']
print("== raw metrics:", rep)
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) for k in keys) and not h.endswith(('.max', '.min')) and 'per_second' not in h.replace('sm__cycles_elapsed.avg.per_second', ''):
        print("  %-90s %-14s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); samp = collections.Counter(); tot = 0
for r in data:
    s = r[ia].strip(); n = int(r[ie]); tot += n
    t = s.split()
    op = (t[1] if s.startswith('@') else t[0]).split('.')[0]
    ops[op] += n; samp[op] += int(r[isamp])
print("== executed warp instructions: %d, static SASS: %d" % (tot, len(data)))
for op, n in ops.most_common(top):
    print("  %-10s %12d %5.1f%%  stall samples %d" % (op, n, 100.0 * n / tot, samp[op]))
