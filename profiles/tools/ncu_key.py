"""Key throughput / stall metrics of the first kernel of an ncu report. usage: ncu_key.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h = rows[0]; r = rows[2]
keys = ['gpu__time_duration.sum', 'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
keys += [k for k in h if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio')]
for k in keys:
    if k in h:
        v = r[h.index(k)]
        try:
            if float(v.replace(',', '')) == 0: continue
        except ValueError: pass
        print(f'{k.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""):75s} {v} {rows[1][h.index(k)]}')
