"""Condense an `ncu --set full` report into the text summary + traffic table committed under profiles/.
usage: python profiles/make_summary.py <report.ncu-rep> <out.txt> [<traffic.json>]
(run in the build container: ncu reads the report without a GPU)."""
import csv
import json
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg.per_second', 'smsp__thread_inst_executed_per_inst_executed.ratio']
STALL = 'smsp__average_warps_issue_stalled_'
ALGO = {'cost_eval': 32768 * 3584 + 512 * 3584, 'sample_gp_tc': 32768 * 3584 * 2 + 2 * 896 * 896 * 4, 'softmax_update': 32768 * 3584,
        'sample_gp_kron_mma': 32768 * 3584 * 2 + 7 * 128 * 128 * 4 + 512 * 3584,
        'sample_gp_kron_umma': 32768 * 3584 * 2 + 2 * 4 * 8 * 224 * 16 * 4 + 512 * 3584,
        'prior_matvec_dof': 2 * 512 * 3584 + 7 * 896 * 4,
        'sample_gp_kron_gen': 32768 * 3584 + 7 * 128 * 128 * 4 + 512 * 3584}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines, traffic = [f'# condensed from {rep} (ncu --set full --clock-control none --import-source on; B200, one launch each)'], {}
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        lines.append(f'\n== {name[:100]}')
        vals = dict(zip(hdr, r))
        for k in KEEP:
            if k in vals:
                lines.append(f'  {k:78s} {units[hdr.index(k)]:14s} {vals[k]}')
        stalls = sorted(((float(v or 0), h[len(STALL):-len('_per_issue_active.ratio')]) for h, v in vals.items()
                         if h.startswith(STALL) and h.endswith('_per_issue_active.ratio')), reverse=True)
        lines.append('  warp stall reasons per issue (top 6): ' + ', '.join(f'{n} {v:.2f}' for v, n in stalls[:6]))
        try:
            rd, wr = float(vals['dram__bytes_read.sum']), float(vals['dram__bytes_write.sum'])
            scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
            rd *= scale[units[hdr.index('dram__bytes_read.sum')]]
            wr *= scale[units[hdr.index('dram__bytes_write.sum')]]
            key = next((k for k in ALGO if k in name), None)
            if key and key not in traffic:
                traffic[key] = dict(dram_read_bytes=rd, dram_write_bytes=wr, algorithmic_bytes=ALGO[key], source=out)
        except Exception:
            pass
    open(out, 'w').write('\n'.join(lines) + '\n')
    if len(sys.argv) > 3:
        try:
            old = json.load(open(sys.argv[3]))
        except Exception:
            old = {}
        old.update(traffic)
        json.dump(old, open(sys.argv[3], 'w'), indent=1)
    print('\n'.join(lines[:4]))


if __name__ == '__main__':
    main()
