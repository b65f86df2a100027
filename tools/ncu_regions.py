"""Summarise an ncu report: headline metrics + executed instructions per SASS region / opcode class.
usage: python tools/ncu_regions.py report.ncu-rep [bin]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, d = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
for w in want:
    if w in hdr:
        print('%-80s %s %s' % (w, d[hdr.index(w)], units[hdr.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
iex, ith, isrc, ism = (hdr.index(k) for k in ('Instructions Executed', 'Avg. Threads Executed', 'Source', '# Samples'))
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[ism]) for r in data)
print('\nSASS instructions: %d static, %d executed (warp-level)' % (len(data), tot))
ops = collections.Counter()
for r in data:
    op = r[isrc].strip().split()
    op = op[1] if op and op[0].startswith('@') else (op[0] if op else '?')
    ops[op.split('.')[0]] += int(r[iex])
print('by opcode:', ', '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in ops.most_common(22)))
print()
for b in range(0, len(data), B):
    seg = data[b:b + B]
    ex = sum(int(r[iex]) for r in seg); sm = sum(int(r[ism]) for r in seg)
    th = sum(float(r[ith] or 0) * int(r[iex]) for r in seg) / max(ex, 1)
    if ex > tot * 0.01:
        print("%4d-%4d  instr %5.1f%%  samples %5.1f%%  avg thr %4.1f   %s" % (b, b + B, 100 * ex / tot, 100 * sm / tots, th,
                                                                              seg[0][isrc].strip()[:50]))
