"""print selected metrics of every launch in an .ncu-rep (ncu --page raw --csv)"""
import csv, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct'] + extra
for r in rows[2:]:
    for k in keys:
        for i, h in enumerate(hdr):
            if h == k or (k.endswith('*') and h.startswith(k[:-1])):
                print('%-90s %s %s' % (h, r[i], units[i]))
    print('-' * 20)
