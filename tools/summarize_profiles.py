"""Turns the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.
  python tools/summarize_profiles.py <launches.csv> <full.ncu-rep> <tag>"""
import collections
import csv
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1:4]
out_dir = os.path.join(REPO, 'profiles')

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ik, iv, iu, ig = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
per = []
for r in rows[1:]:
    v = float(r[iv].replace(',', ''))
    v = v / 1000 if r[iu] == 'ns' else (v * 1000 if r[iu] == 'ms' else v)
    per.append((r[ik], r[ig], v))
agg = collections.OrderedDict()
for name, grid, us in per:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(a[1] for a in agg.values())
with open(os.path.join(out_dir, tag + '_launches_c2_step.csv'), 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 3 (launches of >= 1 step; '
                'cold-cache, serialised: the SHARE per kernel is what matters)'])
    w.writerow(['kernel', 'launches', 'total_us', 'share_pct'])
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([name, a[0], '%.1f' % a[1], '%.2f' % (100 * a[1] / total)])
    w.writerow([])
    w.writerow(['# per launch, in stream order'])
    w.writerow(['kernel', 'grid', 'us'])
    for name, grid, us in per:
        w.writerow([name, grid, '%.2f' % us])

raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u = rr[0], rr[1]
keys = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active']
with open(os.path.join(out_dir, tag + '_ncu_full_conv3x3_rows_summary.csv'), 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['# ncu --set full --clock-control none --import-source on -k regex:conv3x3_rows (consecutive launches of one dense block '
                'inside a bench.py step)'])
    w.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(rr) - 2)])
    for k in keys:
        if k in h:
            i = h.index(k)
            w.writerow([k, u[i]] + [r[i] for r in rr[2:]])
print('wrote', tag)
