"""Loss reductions of the training step as CUDA kernels behind the C-ABI (esr_l1_reduce / esr_l1_grad, esr_bce_rel_loss / _bwd):
the pixel and VGG-feature L1 criteria (models/SRRaGAN_model.py:98,129,434,448-451) and the relativistic average GAN terms
(:353-354 D step, :475-476 G step; GANLoss 'vanilla', models/modules/loss.py:212-246).  CPU tensors (the orchestration tests drive the
model with stand-in networks on the CPU) take torch's own functions."""
import ctypes as C

import torch
import torch.nn as nn

from . import lib as L
from . import parallel


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_ws = {}


def _l1_workspace(dev):
    key = (str(dev), torch.cuda.current_stream().cuda_stream)
    if key not in _ws:
        _ws[key] = torch.empty(int(L.load().esr_l1_workspace_bytes()) // 4, dtype=torch.float32, device=dev)
    return _ws[key]


class _L1MeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.float().contiguous(), b.float().contiguous()
        out = torch.empty((), dtype=torch.float32, device=a.device)
        ws = _l1_workspace(a.device)
        L.check(L.load().esr_l1_reduce(_p(a), _p(b), a.numel(), _p(ws), ws.numel() * 4, _p(out), _stream()))
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.float().contiguous()
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        L.check(L.load().esr_l1_grad(_p(a), _p(b), a.numel(), _p(g), _p(ga), _p(gb), _stream()))
        return ga, gb


def l1_mean(a, b):
    """mean |a - b| (nn.L1Loss); CUDA tensors of equal shape run the two-stage reduction kernel"""
    if a.is_cuda and b.is_cuda and a.shape == b.shape and a.numel() > 0:
        return _L1MeanFn.apply(a, b)
    return torch.nn.functional.l1_loss(a, b)


class L1Loss(nn.L1Loss):
    """nn.L1Loss(reduction='mean') on the esr_l1_reduce kernel"""

    def forward(self, input, target):
        if self.reduction == 'mean':
            return l1_mean(input, target)
        return super(L1Loss, self).forward(input, target)


class _RelBceFn(torch.autograd.Function):
    """(la, lb) = (mean bce(a - mean_global(b), ta), mean bce(b - mean_global(a), tb)) over the LOCAL samples; the means of a and b and
    the derivative sums are taken over the GLOBAL batch (one 2-scalar all-reduce each way when the batch is sharded over ranks, exactly
    what esr_b200.parallel.global_mean_autograd does term by term)."""

    @staticmethod
    def forward(ctx, a, b, ta, tb):
        a32, b32 = a.float().contiguous().view(-1), b.float().contiguous().view(-1)
        n = a32.numel()
        world = parallel.world()
        sums, n_glob = None, float(n)
        if world > 1:
            sums = torch.stack([a32.sum(), b32.sum()])
            torch.distributed.all_reduce(sums)
            n_glob = float(n * world)
        out4 = torch.empty(4, dtype=torch.float32, device=a.device)
        ea, eb = torch.empty_like(a32), torch.empty_like(b32)
        L.check(L.load().esr_bce_rel_loss(_p(a32), _p(b32), n, _p(sums), n_glob, float(ta), float(tb), _p(out4), _p(ea), _p(eb), _stream()))
        ctx.save_for_backward(ea, eb, out4)
        ctx.n, ctx.n_glob, ctx.shape_a, ctx.shape_b = n, n_glob, a.shape, b.shape
        return out4[0] / n, out4[1] / n

    @staticmethod
    def backward(ctx, g_la, g_lb):
        ea, eb, out4 = ctx.saved_tensors
        S = out4[2:4].clone()
        if parallel.world() > 1:
            torch.distributed.all_reduce(S)
        ga = torch.empty_like(ea) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(eb) if ctx.needs_input_grad[1] else None
        g_la = g_la.float().contiguous() if g_la is not None else None
        g_lb = g_lb.float().contiguous() if g_lb is not None else None
        L.check(L.load().esr_bce_rel_loss_bwd(_p(ea), _p(eb), ctx.n, _p(S), ctx.n_glob, _p(g_la), _p(g_lb), _p(ga), _p(gb), _stream()))
        return (ga.view(ctx.shape_a) if ga is not None else None), (gb.view(ctx.shape_b) if gb is not None else None), None, None


def relativistic_bce(pred_a, pred_b, target_a, target_b):
    """(BCEWithLogits(pred_a - mean(pred_b), target_a), BCEWithLogits(pred_b - mean(pred_a), target_b)) with the means over the global
    batch - the two terms of the relativistic average GAN loss, one forward and one backward kernel"""
    return _RelBceFn.apply(pred_a, pred_b, float(target_a), float(target_b))
