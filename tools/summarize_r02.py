"""Turns the scratch ncu outputs of tools/gpu_session_prof.sh (gpurun_out/prof/) into the tracked round-2 summaries under profiles/.
  python tools/summarize_r02.py [gpurun_out/prof]"""
import collections
import csv
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, 'gpurun_out', 'prof')
out_dir = os.path.join(REPO, 'profiles')

KEYS = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active']


def raw_summary(name, dst, title):
    path = os.path.join(src, name)
    if not os.path.exists(path):
        print('missing', path)
        return
    rows = list(csv.reader(open(path)))
    while rows and 'Kernel Name' not in rows[0]:
        rows.pop(0)
    h, u = rows[0], rows[1]
    with open(os.path.join(out_dir, dst), 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['# ' + title])
        w.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(rows) - 2)])
        for k in KEYS:
            if k in h:
                i = h.index(k)
                w.writerow([k, u[i]] + [r[i] for r in rows[2:]])
    print('wrote', dst, len(rows) - 2, 'launches')


def launch_list(name, dst, title, end_marker='adam_multi'):
    path = os.path.join(src, name)
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ik, iv, iu, ig = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('Grid Size')
    per = []
    for r in rows[1:]:
        v = float(r[iv].replace(',', ''))
        v = v / 1000 if r[iu] == 'ns' else (v * 1000 if r[iu] == 'ms' else v)
        per.append((r[ik], r[ig], v))
    ends = [i for i, (k, _, _) in enumerate(per) if end_marker in k]
    a, b = (ends[1] + 1, ends[2] + 1) if len(ends) >= 3 else (0, len(per))     # one steady-state step
    step = per[a:b]
    agg = collections.OrderedDict()
    for k, _, us in step:
        e = agg.setdefault(k, [0, 0.0])
        e[0] += 1
        e[1] += us
    total = sum(e[1] for e in agg.values())
    with open(os.path.join(out_dir, dst), 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['# ' + title])
        w.writerow(['# launches of ONE steady-state step: %d, sum of durations %.1f us (cold-cache, serialised: the SHARE per kernel is what matters)' % (len(step), total)])
        w.writerow(['kernel', 'launches', 'total_us', 'share_pct', 'us_per_launch'])
        for k, e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, e[0], '%.1f' % e[1], '%.2f' % (100 * e[1] / total), '%.2f' % (e[1] / e[0])])
        w.writerow([])
        w.writerow(['# per launch, in stream order'])
        w.writerow(['kernel', 'grid', 'us'])
        for k, g, us in step:
            w.writerow([k, g, '%.2f' % us])
    print('wrote', dst, len(step), 'launches', '%.1f us' % total)


launch_list('launches_train.csv', 'r02_launches_c2_train_step.csv',
            'ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline')
raw_summary('fwd_full.raw.csv', 'r02_ncu_full_conv3x3_rows_summary.csv',
            'ncu --set full --clock-control none -k regex:conv3x3_rows_kernel|conv3x3_wgrad_kernel, 12 consecutive FORWARD launches (dense blocks) of the '
            'third training step of tools/stress_legs.py train --batch 16 --lr 256 (C2 shape, bf16)')
raw_summary('bwd_full.raw.csv', 'r02_ncu_full_conv3x3_bwd_summary.csv',
            'same run, 24 consecutive BACKWARD launches: dgrad (row kernel, epilogues 4 = mask-only slice, 5 = closing launch) interleaved with conv3x3_wgrad_kernel')
raw_summary('cem_full.raw.csv', 'r02_ncu_cem.csv',
            'ncu --set full --clock-control none -k regex:cem_ : cem_down_fast / cem_inv_fast / cem_up_add_fast at C2 (16 x 3 x 1024 x 1024 HR)')
