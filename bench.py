#!/usr/bin/env python
"""bench.py — HR megapixels/s of the RRDB+CEM forward (BASELINE.json configs[1]: RRDBNet nb=23 nf=64 x4,
batch 16 of 256x256 LR per GPU, forward + CEM).  Prints ONE JSON line (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

own arm        : the CUDA path through the reference-facing API (CEM_PyTorch(RRDBNet)); `value` with inputs
                 resident in HBM, `e2e` with pinned-host inputs/outputs copied inside the timed region.
--impl reference: the oracle port of the reference's PyTorch CPU path on the host cores (the reference is pure
                 Python and /root/reference does not exist on the GPU box), a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, 'explorable-super-resolution_b200')
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

NF, NB, SCALE, BATCH, LR = 64, 23, 4, 16, 256
METRIC = 'HR megapixels/sec (fwd) 4x SR RRDB+CEM'
UNIT = 'HR-MP/s'
# conv FLOPs (2*MAC) per HR pixel of RRDBNet nf64 nb23 x4, SURVEY 8(d) / BASELINE.md 4
FLOP_PER_HR_PX = 2.2409e6


def conv_flops_per_lr_px(nf=NF, nb=NB, gc=32, scale=SCALE):
    k = 9 * 2
    trunk = nb * 3 * (sum((nf + i * gc) * gc for i in range(4)) + (nf + 4 * gc) * nf) * k
    lr = (3 * nf + nf * nf) * k
    up, res = 0, 1
    for _ in range(int(round(__import__('math').log2(scale)))):
        res *= 4
        up += nf * nf * k * res
    hr = (nf * nf + nf * 3) * k * res
    return trunk + lr + up + hr


def peaks():
    try:
        with open(os.path.join(REPO, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return p, 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return {'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'hbm_gbs': 6650.0}, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].startswith('Active') for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace('.', '', 1).isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'power_w_max': max(pw) if pw else None, 'samples': len(self.rows)}


def profiled_traffic():
    """DRAM bytes (read + write) of the most expensive conv launch type, from the committed `ncu --set full` summary"""
    import csv
    path = os.path.join(REPO, 'profiles', 'r01b_ncu_full_conv3x3_rows_summary.csv')
    try:
        rows = {r[0]: r for r in csv.reader(open(path)) if r and not r[0].startswith('#')}
        t = [float(v) for v in rows['gpu__time_duration.sum'][2:]]
        k = t.index(max(t))
        mult = lambda u: {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[u]
        rd, wr = rows['dram__bytes_read.sum'], rows['dram__bytes_write.sum']
        return {'bytes': float(rd[2 + k]) * mult(rd[1]) + float(wr[2 + k]) * mult(wr[1]), 'kernel': rows['Kernel Name'][2 + k],
                'launch_us': t[k], 'algorithmic_bytes': 16 * 256 * 256 * (192 + 64 + 64) * 2.0,
                'source': 'profiles/r01b_ncu_full_conv3x3_rows_summary.csv (conv5 192->64 + 16-bit residual, 16x256x256)'}
    except Exception:
        return None


def build_model(dev):
    """CEM_PyTorch(RRDBNet) exactly as models.networks.define_G builds it for training (kaiming x0.1 init, seed 0)."""
    import contextlib
    import io
    import torch
    import models.networks as networks
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    opt = {'gpu_ids': None, 'is_train': True, 'datasets': {'train': {'patch_size': LR * SCALE}},
           'network_G': {'which_model_G': 'RRDB_net', 'latent_input': None, 'latent_input_domain': None, 'in_nc': 3, 'out_nc': 3,
                         'nf': NF, 'nb': NB, 'gc': 32, 'scale': SCALE, 'norm_type': None, 'mode': 'CNA', 'CEM_arch': 1}}
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        cem = CEMnet(Get_CEM_Conf(SCALE))
        net = networks.define_G(opt, CEM=cem, num_latent_channels=0)
    return net.to(dev), cem


def run_reference(args, rank, world, out_fd):
    """oracle port on the host cores; each step = 1 image of the C2 workload (256x256 -> 1024x1024)."""
    if rank != 0:
        return
    import torch
    from oracle import esr_oracle as O
    import models.modules.architecture as arch
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.default_generator.manual_seed(0)     # CPU generator only: this leg must not depend on the device's health
    net = arch.RRDBNet(3, 3, NF, NB, upscale=SCALE, num_latent_channels=0)
    for p in net.parameters():
        if p.dim() > 1:
            torch.nn.init.kaiming_normal_(p, a=0, mode='fan_in')
            p.data *= 0.1
        else:
            p.data.zero_()
    sd = {'generated_image_model.' + k: v.detach() for k, v in net.state_dict().items()}
    cem = CEMnet(Get_CEM_Conf(SCALE))
    x = torch.rand(1, 3, LR, LR)
    step = lambda: O.cem_wrapped_forward(x, sd, cem.ds_kernel, cem.inv_hTh, SCALE, 1, int(cem.invalidity_margins_LR), False, NF, NB)
    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
    mp = (LR * SCALE) ** 2 / 1e6
    val = mp / dt
    sample = '1 of %d images per step (1x3x%dx%d -> %dx%d), fp32, torch CPU, %d threads' % (BATCH, LR, LR, LR * SCALE, LR * SCALE, cores)
    _emit(out_fd, {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'C2: RRDBNet nb=23 nf=64 x4 + CEM forward, batch 16 of 256x256 LR per GPU (reference arm: 1-image sample per step)'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})


def cpu_baseline():
    """oracle port, bounded sample (one 128x128 crop of the C2 workload, scaled per pixel)."""
    import torch
    from oracle import esr_oracle as O
    import models.modules.architecture as arch
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.default_generator.manual_seed(0)     # CPU generator only: this leg must not depend on the device's health
    net = arch.RRDBNet(3, 3, NF, NB, upscale=SCALE, num_latent_channels=0)
    sd = {'generated_image_model.' + k: v.detach() * (0.1 if v.dim() > 1 else 0.0) for k, v in net.state_dict().items()}
    cem = CEMnet(Get_CEM_Conf(SCALE))
    x = torch.rand(1, 3, 128, 128)
    with torch.no_grad():
        O.cem_wrapped_forward(x, sd, cem.ds_kernel, cem.inv_hTh, SCALE, 1, 10, False, NF, NB)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            O.cem_wrapped_forward(x, sd, cem.ds_kernel, cem.inv_hTh, SCALE, 1, 10, False, NF, NB)
        dt = (time.perf_counter() - t0) / reps
    return {'value': (128 * SCALE) ** 2 / 1e6 / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '1x3x128x128 LR crop (1/64 of a step), %d reps, fp32 torch CPU oracle, %d threads' % (reps, cores)}


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: libraries (NCCL prints its version at init) are sent to stderr instead"""
    sys.stdout.flush()
    fd = os.dup(1)
    os.dup2(2, 1)
    return fd


def _emit(fd, obj):
    os.write(fd, (json.dumps(obj) + '\n').encode())


def main():
    out_fd = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='esr_b200')
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the extra generator-training-step measurement')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world, out_fd)
        return

    import torch
    import torch.distributed as dist
    from esr_b200 import lib, ops
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    ops.device_check()
    model, cem = build_model(dev)
    model.train()  # CEM train mode = no padding (config 2: forward + CEM only)
    B = args.batch
    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand(B, 3, LR, LR, generator=gen).pin_memory()
    y_host = torch.empty(B, 3, LR * SCALE, LR * SCALE).pin_memory()
    x_dev = x_host.to(dev)

    # event pairs around every conv launch: the dominant kernel's duration, measured in the timed region
    conv_events = []
    real_conv = ops.conv3x3
    record = {'on': False}

    def timed_conv(*a, **k):
        if not record['on']:
            return real_conv(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_conv(*a, **k)
        e1.record()
        conv_events.append((e0, e1))
    ops.conv3x3 = timed_conv
    import esr_b200.engine as engine_mod
    engine_mod.ops.conv3x3 = timed_conv

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident():
        with torch.no_grad():
            return model(x_dev)

    # e2e: every step copies its input from pinned host memory and its result back to pinned host memory.  The copies run on
    # their own stream, double-buffered, so the D2H of step i overlaps the compute of step i+1 (what a serving loop does);
    # all of them complete inside the timed region (the final synchronize covers the last D2H).
    copy_stream, h2d_stream = torch.cuda.Stream(), torch.cuda.Stream()   # D2H and H2D on their own streams: neither queues behind the other
    y_hosts = [y_host, torch.empty_like(y_host).pin_memory()]
    e2e_state = {'i': 0, 'keep': [None, None]}

    def step_e2e():
        i = e2e_state['i']
        cur = torch.cuda.current_stream()
        with torch.no_grad():
            with torch.cuda.stream(h2d_stream):
                xd = x_host.to(dev, non_blocking=True)
            cur.wait_stream(h2d_stream)
            y = model(xd)
            copy_stream.wait_stream(cur)
            with torch.cuda.stream(copy_stream):
                y_hosts[i & 1].copy_(y, non_blocking=True)
            y.record_stream(copy_stream)
            xd.record_stream(cur)
        e2e_state['keep'][i & 1] = y
        e2e_state['i'] = i + 1
        return y

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # (1) the timed region proper: K steps, nothing but the product's own launches on the stream
    launches0 = lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        step_resident()
    t1.record()
    barrier()
    launches = lib.launch_count() - launches0
    ms = t0.elapsed_time(t1) / args.steps
    # (2) the same K steps again with a CUDA-event pair around every conv launch (the dominant kernel's duration for
    #     the roofline; kept out of (1) because the event records serialise the programmatic dependent launches)
    record['on'] = True
    ops.PLAN_REPLAY = False      # one host call per conv, so that each launch can sit between its own event pair
    barrier()
    for _ in range(args.steps):
        step_resident()
    barrier()
    ops.PLAN_REPLAY = True
    record['on'] = False
    conv_ms = sum(a.elapsed_time(b) for a, b in conv_events) / args.steps
    n_conv = len(conv_events) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    assert lib.watchdog()[0] == 0, 'pipeline watchdog fired'

    # e2e: pinned host -> device, forward, device -> pinned host, every step
    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.current_stream().wait_stream(copy_stream)   # the last result's D2H belongs to the timed region
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    if world > 1:      # max over ranks of the headline numbers, before any extra leg runs
        t = torch.tensor([ms, ms_e2e, conv_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, conv_ms = [float(v) for v in t]

    # generator training step at the same shape (fwd + CEM + L1 + bwd with weight gradients + Adam; bf16 engine), extra key
    train = None
    if not args.no_train:
        try:
            print('[bench] train leg', file=sys.stderr, flush=True)
            from esr_b200 import parallel
            params = [p_ for n_, p_ in model.named_parameters() if 'Filter_OP' not in n_]
            for p_ in params:
                p_.requires_grad_(True)
            opt_g = torch.optim.Adam(params, lr=1e-4)
            hr_dev = torch.rand(B, 3, LR * SCALE, LR * SCALE, device=dev)

            def step_train():
                opt_g.zero_grad(set_to_none=True)
                loss = (model(x_dev) - hr_dev).abs().mean()
                loss.backward()
                parallel.average_gradients(params)
                opt_g.step()
            for _ in range(2):
                step_train()
            barrier()
            l0 = lib.launch_count()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(3):
                step_train()
            g1.record()
            barrier()
            ms_train = g0.elapsed_time(g1) / 3
            if world > 1:
                tt = torch.tensor([ms_train], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms_train = float(tt[0])
            train = {'ms_per_step': ms_train, 'gpu_launches_per_step': (lib.launch_count() - l0) // 3, 'dtype': 'bf16 operands, f32 master weights',
                     'step': 'forward + CEM + L1 loss + backward (dgrad + wgrad) + gradient all-reduce + Adam', 'steps': 3}
            for p_ in params:
                p_.requires_grad_(False)
                p_.grad = None
            assert lib.watchdog()[0] == 0, 'pipeline watchdog fired in the training step'
        except Exception as e:  # the forward numbers above stay valid
            train = {'error': repr(e)[:300]}

    # the CPU baseline and the extra GAN-step leg run after everything the JSON line needs from the device is in host floats
    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            cpu_base = cpu_baseline()
        except Exception as e:
            cpu_base = {'error': repr(e)[:300]}
    # full SRRaGAN step at the per-GPU shape of BASELINE config 3 (batch 4 of 52x52 LR, 208x208 HR patches, 128x128 critic crops):
    # D step (Discriminator_VGG_128, relativistic loss, Adam) + G step (pixel + VGG-feature + relativistic GAN loss, Adam), through
    # create_model / feed_data (host tensors) / optimize_parameters, gradients all-reduced over the ranks.  Extra key.
    gan = None
    if not args.no_train:
        try:
            torch.cuda.synchronize()       # an asynchronous fault of an earlier leg surfaces here, not inside this one
            print('[bench] gan_step leg', file=sys.stderr, flush=True)
            import contextlib, io
            from models import create_model

            class ND(dict):
                def __missing__(self, k):
                    return None
            tr = ND(pixel_weight=1e-2, pixel_criterion='l1', feature_weight=1.0, feature_criterion='l1', gan_type='vanilla', gan_weight=5e-3,
                    lr_G=1e-4, beta1_G=0.9, weight_decay_G=0, lr_D=1e-4, beta1_D=0.9, weight_decay_D=0, D_update_ratio=1, D_init_iters=0,
                    lr_scheme='MultiStepLR', lr_steps=[100000], lr_gamma=0.5, grad_accumulation_steps_G=1, grad_accumulation_steps_D=1)
            o3 = ND(model='srragan', scale=4, gpu_ids=[local], is_train=True, range=[0, 1], train=tr, datasets=ND(train=ND(patch_size=208, batch_size=4)),
                    path=ND(models='/tmp/esr_bench_c3/models', pretrained_model_G=None, log='/tmp/esr_bench_c3'),
                    network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input=None, latent_input_domain=None, latent_channels=None,
                                 norm_type=None, mode='CNA', nf=NF, nb=NB, in_nc=3, out_nc=3, gc=32, scale=4),
                    network_D=ND(which_model_D='discriminator_vgg_128', norm_type='batch', act_type='leakyrelu', mode='CNA', nf=64, in_nc=3))
            with contextlib.redirect_stdout(io.StringIO()):
                m3 = create_model(o3)
            lr3, hr3 = torch.rand(4, 3, 52, 52, generator=gen), torch.rand(4, 3, 208, 208, generator=gen)

            def step_gan():
                m3.feed_data({'LR': lr3, 'HR': hr3})
                m3.optimize_parameters()
            for _ in range(3):
                step_gan()
            barrier()
            l0 = lib.launch_count()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(5):
                step_gan()
            g1.record()
            barrier()
            ms_gan = g0.elapsed_time(g1) / 5
            if world > 1:
                tg = torch.tensor([ms_gan], device=dev)
                dist.all_reduce(tg, op=dist.ReduceOp.MAX)
                ms_gan = float(tg[0])
            gan = {'ms_per_step': ms_gan, 'gpu_launches_per_step': (lib.launch_count() - l0) // 5, 'steps': 5,
                   'config': 'C3 per-GPU shape: batch 4 of 52x52 LR -> 208x208, critic on 128x128 crops, bf16 operands, f32 master weights',
                   'step': 'D step + G step (pixel + VGG-feature + relativistic GAN loss) + gradient all-reduces + two Adam steps',
                   'l_d_real_fake': float(m3.log_dict['l_d_real_fake'][-1][1]), 'l_g_gan': float(m3.log_dict['l_g_gan'][-1][1])}
            del m3
            assert lib.watchdog()[0] == 0, 'pipeline watchdog fired in the GAN step'
        except Exception as e:
            gan = {'error': repr(e)[:300]}

    mp_step = world * B * (LR * SCALE) ** 2 / 1e6
    if rank == 0:
        pk, pk_src = peaks()
        flops_step = conv_flops_per_lr_px() * B * LR * LR  # per GPU
        achieved = flops_step / (conv_ms * 1e-3) / 1e12
        peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
        out = {
            'metric': METRIC, 'value': mp_step / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16 operands, f32 accumulate/trunk',
            'data': 'synthetic',
            'config': {'workload': 'C2: RRDBNet nb=%d nf=%d x%d + CEM forward, batch %d of %dx%d LR per GPU' % (NB, NF, SCALE, B, LR, LR),
                       'weights': 'reference training init (kaiming x0.1), seed 0', 'l2': 'working set per step (>6 GB) far exceeds the 126 MB L2',
                       'parallelism': 'dp%d, batch-sharded, no collective in the forward path' % world},
            'clocks': clocks, 'gpu_launches': launches,
            'e2e': {'value': mp_step / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': x_host.numel() * 4,
                    'd2h_bytes_per_step': y_host.numel() * 4},
            'roofline': {'bound': 'tensor', 'kernel': 'conv3x3_rows_kernel + conv3x3_tc_kernel (%d conv launches/step)' % n_conv, 'achieved': achieved, 'peak': peak,
                         'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': (profiled_traffic() or {}).get('bytes'),
                         'traffic_detail': profiled_traffic(), 'peak_source': pk_src + ', sustained bf16',
                         'kernel_ms_per_step': conv_ms, 'algorithmic_tflop_per_step': flops_step / 1e12},
        }
        if train is not None:
            if 'ms_per_step' in train:
                train['value'] = mp_step / (train['ms_per_step'] * 1e-3)
                train['unit'] = 'HR-MP/s (fwd+bwd)'
            out['train'] = train
        if gan is not None:
            if 'ms_per_step' in gan:
                gan['value'] = world * 4 * 208 * 208 / 1e6 / (gan['ms_per_step'] * 1e-3)
                gan['unit'] = 'HR-MP/s (fwd+bwd, D+G)'
            out['gan_step'] = gan
        if cpu_base is not None:
            out['cpu_baseline'] = cpu_base
        _emit(out_fd, out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
