#!/usr/bin/env python
"""bench.py — HR megapixels/s (fwd+bwd) of the RRDB+CEM generator step at BASELINE.json configs[1]'s shape (C2: RRDBNet nb=23 nf=64
x4, batch 16 of 256x256 LR per GPU).  Prints ONE JSON line (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

own arm         : one step = forward + CEM projection + L1 loss + backward (dgrad + wgrad) + gradient all-reduce (N > 1) + Adam.
                  `value` with the batch resident in HBM; `e2e` the same step through the reference-facing API
                  (create_model / feed_data(host batch) / optimize_parameters) with the host->device copies of LR and HR and the
                  device->host read of the logged loss inside the timed region.  Extra keys (never the headline): `forward` (the
                  forward + CEM only numbers of configs[1]), `gan_step` (C3, full SRRaGAN step), `zopt` (C4, Z_optimizer loop).
--impl reference: the oracle port of the reference's PyTorch CPU path on the host cores, forward + backward + Adam on a bounded
                  sample per step (the reference is pure Python; /root/reference does not exist on the GPU box).

Robustness: the headline is complete (host floats) before any extra leg runs; every extra leg is contained; a failing rank raises
an abort flag in the rendezvous store which a watcher thread on every rank turns into "rank 0 prints what it has, everyone
os._exit()s" instead of a collective hang; the process group has a 120 s timeout.
"""
import argparse
import datetime
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
METRIC = 'HR megapixels/sec (fwd+bwd) 4x SR RRDB+CEM'
UNIT = 'HR-MP/s'
WORKLOAD = 'C2: RRDBNet nb=%d nf=%d x%d + CEM, batch %d of %dx%d LR per GPU, forward + backward + Adam (L1 loss)'


def conv_flops_per_lr_px(nf=NF, nb=NB, gc=32, scale=SCALE):
    """conv FLOPs (2*MAC) of one RRDBNet forward per LR pixel (36.714 GFLOP per 1024 LR px at nf64/nb23/x4)"""
    import math
    k = 9 * 2
    trunk = nb * 3 * (sum((nf + i * gc) * gc for i in range(4)) + (nf + 4 * gc) * nf) * k
    lr = (3 * nf + nf * nf) * k
    up, res = 0, 1
    for _ in range(int(round(math.log2(scale)))):
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


def _first(d, keys, default=None):
    for k in keys:
        if isinstance(d, dict) and d.get(k) is not None:
            return d[k]
    return default


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
    for name in ('r02_ncu_full_conv3x3_rows_summary.csv', 'r01b_ncu_full_conv3x3_rows_summary.csv'):
        path = os.path.join(REPO, 'profiles', name)
        try:
            rows = {r[0]: r for r in csv.reader(open(path)) if r and not r[0].startswith('#')}
            t = [float(v) for v in rows['gpu__time_duration.sum'][2:]]
            names = rows['Kernel Name'][2:]
            # the first conv5 launch of a dense block (epilogue 2: 192 -> 64 + 16-bit residual; later ones also carry the fp32 trunk)
            k = next((i for i, nm in enumerate(names) if ', 0, 2>' in nm), t.index(max(t)))
            mult = lambda u: {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[u]
            rd, wr = rows['dram__bytes_read.sum'], rows['dram__bytes_write.sum']
            return {'bytes': float(rd[2 + k]) * mult(rd[1]) + float(wr[2 + k]) * mult(wr[1]), 'kernel': rows['Kernel Name'][2 + k],
                    'launch_us': t[k], 'algorithmic_bytes': 16 * 256 * 256 * (192 + 64 + 64) * 2.0,
                    'source': 'profiles/%s (conv5 192->64 + 16-bit residual, 16x256x256)' % name}
        except Exception:
            continue
    return None


class ND(dict):
    def __missing__(self, k):
        return None


def model_options(local, nb=NB, patch=LR * SCALE, batch=BATCH, gan=False, tag='c2'):
    """option tree of the reference's create_model (options/train/train_esrgan.json with CEM_arch) for the benchmarked steps"""
    tr = ND(pixel_weight=1.0, pixel_criterion='l1', lr_G=1e-4, beta1_G=0.9, weight_decay_G=0, lr_scheme='MultiStepLR', lr_steps=[100000],
            lr_gamma=0.5, grad_accumulation_steps_G=1, grad_accumulation_steps_D=1)
    if gan:
        tr.update(pixel_weight=1e-2, feature_weight=1.0, feature_criterion='l1', gan_type='vanilla', gan_weight=5e-3, lr_D=1e-4, beta1_D=0.9,
                  weight_decay_D=0, D_update_ratio=1, D_init_iters=0)
    o = ND(model='srragan', scale=SCALE, gpu_ids=[local], is_train=True, range=[0, 1], train=tr,
           datasets=ND(train=ND(patch_size=patch, batch_size=batch)),
           path=ND(models='/tmp/esr_bench_%s/models' % tag, pretrained_model_G=None, log='/tmp/esr_bench_%s' % tag),
           network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input=None, latent_input_domain=None, latent_channels=None,
                        norm_type=None, mode='CNA', nf=NF, nb=nb, in_nc=3, out_nc=3, gc=32, scale=SCALE))
    if gan:
        o['network_D'] = ND(which_model_D='discriminator_vgg_128', norm_type='batch', act_type='leakyrelu', mode='CNA', nf=64, in_nc=3)
    return o


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


# ------------------------------------------------------------------------------------------------ reference arm / CPU baseline
def _oracle_train_step(lr_px, threads):
    """(step function, HR megapixels per step): oracle restatement of the reference generator, fwd + CEM + L1 + bwd + Adam on the
    host cores for ONE lr_px x lr_px image (fp32, torch CPU)"""
    import torch
    from oracle import esr_oracle as O
    import models.modules.architecture as arch
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    torch.set_num_threads(threads)
    torch.default_generator.manual_seed(0)     # CPU generator only: this leg must not depend on the device's health
    net = arch.RRDBNet(3, 3, NF, NB, upscale=SCALE, num_latent_channels=0)
    sd = {}
    for k, v in net.state_dict().items():
        t = v.detach().clone().float()
        if t.dim() > 1:
            torch.nn.init.kaiming_normal_(t, a=0, mode='fan_in')
            t *= 0.1
        else:
            t.zero_()
        sd['generated_image_model.' + k] = t.requires_grad_(True)
    cem = CEMnet(Get_CEM_Conf(SCALE))
    x = torch.rand(1, 3, lr_px, lr_px)
    hr = torch.rand(1, 3, lr_px * SCALE, lr_px * SCALE)
    opt = torch.optim.Adam(list(sd.values()), lr=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        out = O.cem_wrapped_forward(x, sd, cem.ds_kernel, cem.inv_hTh, SCALE, 1, int(cem.invalidity_margins_LR), False, NF, NB)
        loss = (out - hr).abs().mean()
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step, (lr_px * SCALE) ** 2 / 1e6


def run_reference(args, rank, world, out_fd):
    """reference arm: fwd+bwd+Adam of the oracle port; each step = one 128x128 LR crop (1/64 of a per-GPU C2 step)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    px = 128
    step, mp = _oracle_train_step(px, cores)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = mp / dt
    sample = '1 image of %dx%d LR -> %dx%d per step (1/%d of a per-GPU step), forward + CEM + L1 + backward + Adam, fp32, torch CPU, %d threads' % (
        px, px, px * SCALE, px * SCALE, BATCH * (LR // px) ** 2, cores)
    _emit(out_fd, {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': max(args.warmup, 1),
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': (WORKLOAD % (NB, NF, SCALE, BATCH, LR, LR)) + ' (reference arm: bounded sample per step, rank 0 only)'},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})


def cpu_baseline():
    """oracle port, bounded sample (one 128x128 crop of the C2 workload, fwd+bwd+Adam, scaled per pixel)."""
    cores = os.cpu_count() or 1
    step, mp = _oracle_train_step(128, cores)
    step()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return {'value': mp / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '1x3x128x128 LR crop (1/64 of a step), forward + CEM + L1 + backward + Adam, %d reps, fp32 torch CPU oracle, %d threads' % (reps, cores)}


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: libraries (NCCL prints its version at init) are sent to stderr instead"""
    sys.stdout.flush()
    fd = os.dup(1)
    os.dup2(2, 1)
    return fd


def _emit(fd, obj):
    os.write(fd, (json.dumps(obj) + '\n').encode())


def log(*a):
    print('[bench]', *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ own arm
def supervise(rank, world, out_fd, max_attempts=3):
    """Every rank runs the benchmark in a CHILD process (fresh CUDA context) and keeps no GPU state itself.  A rare device exception
    ('unspecified launch failure', DESIGN.md section 5: seen in the first backward passes of a fresh process, never in steady state) is
    sticky for its process: when rank 0's child ends without having printed the line, every rank starts a fresh child - at most
    twice, and the line says so ('attempt').  The ranks' parents agree through a small TCP store (MASTER_PORT + 1717; the children
    rendezvous on MASTER_PORT + 1718 + attempt, away from the ports a driver may use for its next run)."""
    import subprocess
    base = int(os.environ.get('MASTER_PORT', '29500'))
    addr = os.environ.get('MASTER_ADDR', '127.0.0.1')
    store = None
    if world > 1:
        import datetime as _dt
        from torch.distributed import TCPStore
        store = TCPStore(addr, base + 1717, world, is_master=(rank == 0), timeout=_dt.timedelta(seconds=1800))
    rc = 1
    for a in range(max_attempts):
        env = dict(os.environ, ESR_BENCH_WORKER='1', ESR_BENCH_ATTEMPT=str(a), MASTER_ADDR=addr, MASTER_PORT=str(base + 1718 + a))
        env.pop('TORCHELASTIC_USE_AGENT_STORE', None)      # the children rendezvous on their own store (fresh port per attempt)
        rc = subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, stdout=out_fd)
        if world > 1:
            key = 'esr_bench_attempt_%d' % a
            try:
                if rank == 0:
                    store.set(key, str(rc))
                    rc0 = rc
                    t0 = time.time()      # the store lives in this process: stay until every other rank has read the verdict
                    while store.add(key + '_ack', 0) < world - 1 and time.time() - t0 < 300:
                        time.sleep(0.05)
                else:
                    rc0 = int(store.get(key).decode())
                    store.add(key + '_ack', 1)
            except Exception as e:      # the store's host is gone: act on the own child's result
                log('rank %d: supervisor store: %r' % (rank, e))
                rc0 = rc
        else:
            rc0 = rc
        if rc0 == 0:
            return 0 if rank == 0 else rc
        log('rank %d: attempt %d ended without a result (rank 0 exit code %d)%s' % (rank, a + 1, rc0, '; starting a fresh process' if a + 1 < max_attempts else ''))
    return rc


def main():
    out_fd = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='esr_b200')
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the extra legs (forward-only, C3 SRRaGAN step, C4 Z-optimisation)')
    ap.add_argument('--no-train', action='store_true', help=argparse.SUPPRESS)      # round-1 spelling of --no-extras
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world, out_fd)
        return
    if os.environ.get('ESR_BENCH_WORKER') != '1' and os.environ.get('ESR_BENCH_NO_SUPERVISOR') != '1':
        sys.exit(supervise(rank, world, out_fd))
    attempt = int(os.environ.get('ESR_BENCH_ATTEMPT', '0'))
    warmup = max(args.warmup, 3)
    os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '1')
    os.environ.setdefault('CUDA_MODULE_LOADING', 'EAGER')      # no kernel-image loading while other kernels are in flight

    import torch
    import torch.distributed as dist
    from esr_b200 import lib, ops, parallel
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)

    # ---- fail-fast plumbing: headline line kept as soon as it exists; any rank's failure ends every rank promptly
    state = {'line': None, 'emitted': False, 'lock': threading.Lock(), 'store': None}

    def emit_and_exit(code, note=None):
        with state['lock']:
            if rank == 0 and state['line'] is not None and not state['emitted']:
                if note:
                    state['line'].setdefault('notes', []).append(note)
                _emit(out_fd, state['line'])
                state['emitted'] = True
        os._exit(code if (state['line'] is None or rank != 0) else 0)

    def raise_abort(why):
        log('rank %d aborts: %s' % (rank, why))
        try:
            if state['store'] is not None:
                state['store'].set('esr_bench_abort', ('rank %d: %s' % (rank, why))[:200])
        except Exception:
            pass

    def watcher():
        while True:
            time.sleep(1.0)
            try:
                if state['store'] is not None and state['store'].check(['esr_bench_abort']):
                    why = state['store'].get('esr_bench_abort').decode(errors='replace')
                    log('rank %d leaves on the abort flag (%s)' % (rank, why))
                    emit_and_exit(1, 'aborted: ' + why)
            except Exception:
                pass

    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=120))
        try:
            state['store'] = dist.distributed_c10d._get_default_store()
            threading.Thread(target=watcher, daemon=True).start()
        except Exception:
            state['store'] = None
    ops.device_check()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor([float(v) for v in vals], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    B = args.batch
    HRP = LR * SCALE
    gen = torch.Generator().manual_seed(1234 + rank)
    lr_host = torch.rand(B, 3, LR, LR, generator=gen).pin_memory()
    hr_host = torch.rand(B, 3, HRP, HRP, generator=gen).pin_memory()
    mp_step = world * B * HRP ** 2 / 1e6
    flops_fwd = conv_flops_per_lr_px() * B * LR * LR        # per GPU

    try:
        # =================================================================== headline: generator training step, batch resident
        model, cem = build_model(dev)
        model.train()
        params = [p_ for n_, p_ in model.named_parameters() if 'Filter_OP' not in n_]
        for p_ in params:
            p_.requires_grad_(True)
        if world > 1:
            parallel.broadcast_parameters(model)
        from esr_b200.losses import l1_mean
        from esr_b200.optim import FlatAdam
        opt_g = FlatAdam(params, lr=1e-4).register()      # torch.optim.Adam's arithmetic, one esr_adam_multi launch per step
        x_dev, hr_dev = lr_host.to(dev), hr_host.to(dev)

        def step_train():
            opt_g.zero_grad(set_to_none=True)
            loss = l1_mean(model(x_dev), hr_dev)            # nn.L1Loss on the esr_l1_reduce / esr_l1_grad kernels
            loss.backward()
            parallel.average_gradients(params, optimizer=opt_g)
            opt_g.step()
            return loss

        for _ in range(warmup):
            step_train()
        barrier()
        assert lib.watchdog()[0] == 0, 'pipeline watchdog fired during warm-up'
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = lib.launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record()
        for _ in range(args.steps):
            loss = step_train()
        t1.record()
        barrier()
        launches = lib.launch_count() - launches0
        ms = t0.elapsed_time(t1) / args.steps
        loss_end = float(loss)
        clocks = sampler.stop() if rank == 0 else None
        assert lib.watchdog()[0] == 0, 'pipeline watchdog fired in the timed region'
        assert loss_end == loss_end, 'training loss is NaN'

        # ---- the same step with a CUDA-event pair around every tensor-core launch and every CEM launch (kept out of the timed
        #      region above: event records serialise the programmatic dependent launches)
        ev = {'fwd': [], 'dgrad': [], 'wgrad': [], 'cem': []}
        real = {k: getattr(ops, k) for k in ('conv3x3', 'conv3x3_wgrad', 'cem_down', 'cem_inv', 'cem_up_add', 'sep_adjoint_2d')}

        def timed(kind_of, fn):
            def f(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(*a, **k)
                e1.record()
                ev[kind_of(a, k)].append((e0, e1))
                return r
            return f
        is_bwd = lambda a, k: 'dgrad' if (a[1].transpose_flip or k.get('mask16') is not None or k.get('res3') is not None) else 'fwd'
        ops.conv3x3 = timed(is_bwd, real['conv3x3'])
        ops.conv3x3_wgrad = timed(lambda a, k: 'wgrad', real['conv3x3_wgrad'])
        for name in ('cem_down', 'cem_inv', 'cem_up_add', 'sep_adjoint_2d'):
            setattr(ops, name, timed(lambda a, k: 'cem', real[name]))
        ops.PLAN_REPLAY = False      # one host call per conv, so that each launch sits between its own event pair
        ev_steps = min(args.steps, 3)
        barrier()
        for _ in range(ev_steps):
            step_train()
        barrier()
        ops.PLAN_REPLAY = True
        for name, fn in real.items():
            setattr(ops, name, fn)
        kms = {k: sum(a.elapsed_time(b) for a, b in v) / ev_steps for k, v in ev.items()}
        kcount = {k: len(v) // ev_steps for k, v in ev.items()}
        tc_ms = kms['fwd'] + kms['dgrad'] + kms['wgrad']

        # =================================================================== e2e: the reference-facing API with host batches
        import contextlib
        import io
        from models import create_model
        del model, opt_g, params, x_dev, hr_dev
        torch.cuda.empty_cache()
        with contextlib.redirect_stdout(io.StringIO()):
            m2 = create_model(model_options(local))
        batch = {'LR': lr_host, 'HR': hr_host}

        def step_e2e():
            m2.feed_data(batch)            # pinned host -> device copies of LR and HR (SRRaGAN_model.feed_data)
            m2.optimize_parameters()       # fwd + CEM + crop + L1 + bwd + all-reduce + Adam; the logged loss comes back through a pinned buffer
                                           # (device->host copy every step, appended to log_dict when it has arrived / when the log is read)
        for _ in range(max(warmup, 3)):    # (the model's first iteration is the reference's idle one: SRRaGAN_model.py:351)
            step_e2e()
        barrier()
        n_logged0 = len(m2.log_dict['l_g_pix'])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_e2e()
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1) / args.steps
        assert len(m2.log_dict['l_g_pix']) - n_logged0 == args.steps, 'the e2e steps did not all run a generator update'
        d2h = 4 * 1      # the logged pixel loss (one fp32 scalar per step)
        del m2
        torch.cuda.empty_cache()
        ms, ms_e2e, tc_ms = max_over_ranks([ms, ms_e2e, tc_ms])
    except BaseException as e:  # noqa: BLE001   no headline, no number: leave at once, on every rank
        import traceback
        traceback.print_exc()
        raise_abort('headline leg failed: %r' % (e,))
        os._exit(1)

    pk, pk_src = peaks()
    peak_tf = _first(pk, ['bf16_tflops_sustained', 'bf16_tflops'], 1400.0)
    peak_bw = _first(pk, ['hbm_gbs_sustained', 'hbm_gbs', 'hbm_gb_s'], 6650.0)
    achieved = 3 * flops_fwd / (tc_ms * 1e-3) / 1e12
    cem_bytes = B * 3 * 4.0 * (2 * (HRP * HRP * 2 + LR * LR * 4))     # projection (G, out, x, e, f) + its adjoint, fp32
    traffic = profiled_traffic()
    line = {
        'metric': METRIC, 'value': mp_step / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16 operands, f32 accumulate / master weights / gradients', 'data': 'synthetic',
        'config': {'workload': WORKLOAD % (NB, NF, SCALE, B, LR, LR), 'weights': 'reference training init (kaiming x0.1), seed 0',
                   'l2': 'working set per step (>30 GB of saved activations) far exceeds the 126 MB L2',
                   'parallelism': 'dp%d, batch-sharded, one NCCL all-reduce of the gradients per step' % world},
        'clocks': clocks, 'gpu_launches': launches, 'loss_after_timed_steps': loss_end,
        'e2e': {'value': mp_step / (ms_e2e * 1e-3), 'unit': UNIT, 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': (lr_host.numel() + hr_host.numel()) * 4, 'd2h_bytes_per_step': d2h,
                'api': 'models.create_model(opt) -> feed_data({LR, HR} pinned host tensors) -> optimize_parameters()'},
        'roofline': {'bound': 'tensor', 'kernel': 'conv3x3_rows_kernel (forward + dgrad) and conv3x3_wgrad_kernel: %d + %d + %d launches/step'
                     % (kcount['fwd'], kcount['dgrad'], kcount['wgrad']),
                     'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                     'traffic': (traffic or {}).get('bytes'), 'traffic_detail': traffic, 'peak_source': pk_src + ', sustained bf16',
                     'kernel_ms_per_step': tc_ms, 'algorithmic_tflop_per_step': 3 * flops_fwd / 1e12,
                     'by_kernel': {k: {'ms_per_step': kms[k], 'launches': kcount[k], 'tflops': flops_fwd / (kms[k] * 1e-3) / 1e12,
                                       'frac': flops_fwd / (kms[k] * 1e-3) / 1e12 / peak_tf} for k in ('fwd', 'dgrad', 'wgrad') if kms[k] > 0},
                     'hbm_kernels': {'cem_projection_fwd_bwd': {'ms_per_step': kms['cem'], 'launches': kcount['cem'], 'algorithmic_bytes': cem_bytes,
                                                                'achieved_gbs': cem_bytes / (kms['cem'] * 1e-3) / 1e9 if kms['cem'] > 0 else None,
                                                                'peak_gbs': peak_bw}}},
    }
    if attempt:
        line['attempt'] = attempt + 1
        line.setdefault('notes', []).append('attempt %d: the earlier process(es) ended on a device exception before the headline existed' % (attempt + 1))
    state['line'] = line

    # =================================================================== extra legs (contained; never touch the headline)
    def extra(name, fn, deadline_s):
        if args.no_extras or args.no_train:
            return
        log('extra leg:', name)
        timer = threading.Timer(deadline_s, lambda: (raise_abort('%s exceeded %d s' % (name, deadline_s)), emit_and_exit(1, name + ' timed out')))
        timer.daemon = True
        timer.start()
        try:
            torch.cuda.synchronize()
            line[name] = fn()
            torch.cuda.synchronize()
            assert lib.watchdog()[0] == 0, 'pipeline watchdog fired'
        except BaseException as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            line[name] = {'error': repr(e)[:300]}
            if world > 1:       # the other ranks may sit in a collective of this leg: end the run everywhere, headline kept
                raise_abort('%s failed: %r' % (name, e))
                emit_and_exit(1, '%s failed on rank %d' % (name, rank))
        finally:
            timer.cancel()

    def leg_forward():
        """configs[1] as written: forward + CEM only (the round-1 headline), resident and with host buffers"""
        model, _ = build_model(dev)
        model.train()
        x_dev = lr_host.to(dev)
        y_hosts = [torch.empty(B, 3, HRP, HRP).pin_memory() for _ in range(2)]

        def fwd():
            with torch.no_grad():
                return model(x_dev)
        for _ in range(warmup):
            fwd()
        barrier()
        l0 = lib.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fwd()
        b.record()
        barrier()
        ms_f = a.elapsed_time(b) / args.steps
        nl = (lib.launch_count() - l0) // args.steps
        conv_ev, cem_ev = [], []
        real_ops = {k: getattr(ops, k) for k in ('conv3x3', 'cem_down', 'cem_inv', 'cem_up_add')}

        def timed_into(lst, fn):
            def f(*aa, **kk):
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                r = fn(*aa, **kk)
                c1.record()
                lst.append((c0, c1))
                return r
            return f
        ops.conv3x3, ops.PLAN_REPLAY = timed_into(conv_ev, real_ops['conv3x3']), False
        for k in ('cem_down', 'cem_inv', 'cem_up_add'):
            setattr(ops, k, timed_into(cem_ev, real_ops[k]))
        try:
            for _ in range(3):
                fwd()
            barrier()
        finally:
            ops.PLAN_REPLAY = True
            for k, fn in real_ops.items():
                setattr(ops, k, fn)
        conv_ms = sum(c0.elapsed_time(c1) for c0, c1 in conv_ev) / 3
        cem_ms = sum(c0.elapsed_time(c1) for c0, c1 in cem_ev) / 3
        cem_alg = B * 3 * 4.0 * (2 * HRP * HRP + LR * LR)         # read G, read x, write out: 24.75 B per HR pixel
        copy_stream, h2d_stream = torch.cuda.Stream(), torch.cuda.Stream()
        keep = [None, None]

        def fwd_e2e(i):
            cur = torch.cuda.current_stream()
            with torch.no_grad():
                with torch.cuda.stream(h2d_stream):
                    xd = lr_host.to(dev, non_blocking=True)
                cur.wait_stream(h2d_stream)
                y = model(xd)
                copy_stream.wait_stream(cur)
                with torch.cuda.stream(copy_stream):
                    y_hosts[i & 1].copy_(y, non_blocking=True)
                y.record_stream(copy_stream)
                xd.record_stream(cur)
            keep[i & 1] = y
        for i in range(2):
            fwd_e2e(i)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(args.steps):
            fwd_e2e(i)
        torch.cuda.current_stream().wait_stream(copy_stream)
        b.record()
        barrier()
        ms_fe = a.elapsed_time(b) / args.steps
        ms_f, ms_fe, conv_ms = max_over_ranks([ms_f, ms_fe, conv_ms])
        ach = flops_fwd / (conv_ms * 1e-3) / 1e12
        return {'metric': 'HR megapixels/sec (fwd) 4x SR RRDB+CEM', 'value': mp_step / (ms_f * 1e-3), 'unit': UNIT, 'ms_per_step': ms_f,
                'gpu_launches_per_step': nl, 'dtype': 'f16 operands, f32 accumulate/trunk',
                'e2e': {'value': mp_step / (ms_fe * 1e-3), 'ms_per_step': ms_fe, 'h2d_bytes_per_step': lr_host.numel() * 4,
                        'd2h_bytes_per_step': y_hosts[0].numel() * 4},
                'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach / peak_tf, 'kernel_ms_per_step': conv_ms},
                'cem_projection': {'bound': 'hbm', 'kernels': 'cem_down_fast + cem_inv_fast + cem_up_add_fast', 'ms_per_step': cem_ms,
                                   'launches': len(cem_ev) // 3, 'algorithmic_bytes': cem_alg, 'achieved_gbs': cem_alg / (cem_ms * 1e-3) / 1e9,
                                   'peak_gbs': peak_bw, 'frac': cem_alg / (cem_ms * 1e-3) / 1e9 / peak_bw}}

    def leg_gan():
        """C3: full SRRaGAN step at the per-GPU shape (batch 4 of 52x52 LR -> 208x208, critic and VGG on 128x128 crops)"""
        import contextlib
        import io
        from models import create_model
        with contextlib.redirect_stdout(io.StringIO()):
            m3 = create_model(model_options(local, patch=208, batch=4, gan=True, tag='c3'))
        lr3, hr3 = torch.rand(4, 3, 52, 52, generator=gen).pin_memory(), torch.rand(4, 3, 208, 208, generator=gen).pin_memory()

        def step_gan():
            m3.feed_data({'LR': lr3, 'HR': hr3})
            m3.optimize_parameters()
        for _ in range(4):
            step_gan()
        barrier()
        l0 = lib.launch_count()
        n = 10
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            step_gan()
        b.record()
        barrier()
        ms_gan = max_over_ranks([a.elapsed_time(b) / n])[0]
        # per sample: G 3x fwd, D step 2 images x 3 x 4.454 GFLOP, G-side critic 13.4, VGG19 38.2
        flop = 4 * (3 * conv_flops_per_lr_px() * 52 * 52 + (26.7 + 13.4 + 38.2) * 1e9)
        tf = flop / (ms_gan * 1e-3) / 1e12
        return {'config': 'C3 per-GPU shape: batch 4 of 52x52 LR -> 208x208, critic + VGG19 on 128x128 crops, bf16 operands, f32 master weights',
                'step': 'D step + G step (pixel + VGG-feature + relativistic GAN loss) + gradient all-reduces + two Adam steps',
                'ms_per_step': ms_gan, 'gpu_launches_per_step': (lib.launch_count() - l0) // n, 'steps': n,
                'value': world * 4 * 208 * 208 / 1e6 / (ms_gan * 1e-3), 'unit': 'HR-MP/s (fwd+bwd, D+G)',
                'roofline': {'bound': 'tensor (launch-latency bound in practice)', 'achieved': tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': tf / peak_tf,
                             'algorithmic_tflop_per_step': flop / 1e12},
                'l_d_real_fake': float(m3.log_dict['l_d_real_fake'][-1][1]), 'l_g_gan': float(m3.log_dict['l_g_gan'][-1][1])}

    def leg_zopt():
        """C4: Z_optimizer, objective l1, 100 iterations over 8 regions of 64x64 LR (eval-mode CEM pads to 84x84); replicas only"""
        import contextlib
        import io
        import numpy as np
        from models import create_model
        from Z_optimization import Z_optimizer
        o4 = ND(model='srragan', scale=4, gpu_ids=[local], is_train=False, range=[0, 1], path=ND(pretrained_model_G=None),
                network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input='all_layers', latent_input_domain='HR_downscaled',
                             latent_channels='SVDinNormedOut_structure_tensor', norm_type=None, mode='CNA', nf=NF, nb=NB, in_nc=3, out_nc=3, gc=32, scale=4))
        with contextlib.redirect_stdout(io.StringIO()):
            m4 = create_model(o4)
        regions, iters = 8, 100
        data = {'LR': torch.rand(regions, 3, 64, 64, generator=gen).to(dev), 'desired': torch.rand(regions, 3, 256, 256, generator=gen).to(dev)}
        m4.feed_data({'LR': data['LR'], 'Z': 0}, need_GT=False)      # what the GUI does before it builds a Z_optimizer (GUI.py:1689-1702)
        m4.test()

        def once():
            with contextlib.redirect_stdout(io.StringIO()):
                zo = Z_optimizer(objective='l1', Z_size=[256, 256], model=m4, Z_range=1, max_iters=iters, data=data, initial_LR=0.1,
                                 batch_size=regions, loggers=None)
                zo.optimize()
            return zo
        once()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        zo = once()
        b.record()
        torch.cuda.synchronize()
        sec = max_over_ranks([a.elapsed_time(b) * 1e-3])[0]
        flop = iters * 2 * conv_flops_per_lr_px() * (2.2896 / 2.2409) * regions * 84 * 84
        tf = flop / sec / 1e12
        return {'config': 'C4: Z_optimizer l1, %d iterations x %d regions of 64x64 LR (padded to 84x84), latent model, one GPU (replicas only)' % (iters, regions),
                's_per_%d_iters' % iters: sec, 'value': world * regions * 256 * 256 * iters / 1e6 / sec, 'unit': 'HR-MP/s (fwd + dgrad per iteration)',
                'roofline': {'bound': 'tensor (launch-latency bound in practice)', 'achieved': tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': tf / peak_tf,
                             'algorithmic_tflop': flop / 1e12},
                'final_loss': float(np.asarray(zo.loss_values[-1])) if getattr(zo, 'loss_values', None) else None}

    extra('forward', leg_forward, 240)
    extra('gan_step', leg_gan, 240)
    extra('zopt', leg_zopt, 240)
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            line['cpu_baseline'] = cpu_baseline()
        except Exception as e:
            line['cpu_baseline'] = {'error': repr(e)[:300]}
    with state['lock']:
        if rank == 0 and not state['emitted']:
            _emit(out_fd, line)
            state['emitted'] = True
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass
    os._exit(0)


if __name__ == '__main__':
    main()
