"""print selected metrics of every launch in an `ncu --page raw --csv` dump (the .ncu-rep itself stays on the GPU box)
  python tools/ncu_raw_keys.py <raw.csv> [extra metric names / prefixes*]"""
import csv
import sys

KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
        'sm__inst_executed_pipe_tensor.sum', 'sm__inst_executed_pipe_uniform.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct', 'sm__inst_executed.avg.per_cycle_elapsed']


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    while rows and 'Kernel Name' not in rows[0]:
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    keys = KEYS + sys.argv[2:]
    for r in rows[2:]:
        for k in keys:
            for i, h in enumerate(hdr):
                if h == k or (k.endswith('*') and h.startswith(k[:-1])):
                    print('%-90s %s %s' % (h, r[i], units[i]))
        print('-' * 20)


if __name__ == '__main__':
    main()
