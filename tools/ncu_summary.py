"""print a short summary of an .ncu-rep (raw page) -- the metrics DESIGN.md / profiles/ quote"""
import csv
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg',
    'sm__cycles_active.max', 'sm__cycles_active.min',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.min.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.max.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum',
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        name = v[h.index('Kernel Name')] if 'Kernel Name' in h else ''
        print('kernel:', name[:100])
        for w in WANT:
            if w in h:
                i = h.index(w)
                print('  %-78s %-10s %s' % (w, u[i], v[i]))
        stalls = [(float(v[i].replace(',', '')), n) for i, n in enumerate(h)
                  if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('_per_issue_active.ratio') and v[i]]
        stalls.sort(reverse=True)
        for s, n in stalls[:8]:
            print('  stall %-72s %.3f' % (n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), s))


if __name__ == '__main__':
    main(sys.argv[1])
