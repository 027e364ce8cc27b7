import csv, sys, subprocess, collections, re
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct',
        'sm__inst_executed_pipe_fma.sum.pct', 'sm__inst_executed_pipe_alu.sum.pct', 'sm__inst_executed_pipe_lsu.sum.pct', 'sm__inst_executed_pipe_xu.sum.pct',
        'sm__pipe_fma_cycles_active.avg.pct', 'sm__pipe_tensor', 'sm__cycles_elapsed.avg ', 'launch__waves', 'issue_stalled', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct', 'gpu__dram_throughput', 'l1tex__data_bank_conflicts', 'smsp__thread_inst_executed_per_inst', 'lts__t_bytes.sum ', 'launch__shared_mem', 'sm__sass_thread_inst_executed_op_f']
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:80])
    for h, u, v in zip(hdr, units, r):
        if any(w in h for w in want) and 'peak_sustained_elapsed' not in h.replace('sm__throughput.avg.pct_of_peak_sustained_elapsed', '') and 'per_second' not in h:
            print(f'  {h:85s} {u:12s} {v}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r][0]
hdr = rows[hi]
ci, si, sm = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
ops = collections.Counter(); smp = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    try: n = int(r[ci]); s = int(r[sm])
    except Exception: continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si].strip())
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += n; smp[op] += s; tot += n
print('  static SASS instructions:', len(rows) - hi - 1, ' executed warp instructions:', tot)
for op, n in ops.most_common(14):
    print(f'   {op:10s} {n / tot * 100:5.1f}%   stall samples {smp[op]}')
