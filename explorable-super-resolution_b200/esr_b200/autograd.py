"""Gradient path of the generator (+CEM) w.r.t. its INPUT — what the latent-exploration loop needs
(Z_optimization.py:673-749: loss.backward() through a frozen G down to the latent map Z).

torch.autograd only sees one node: the forward runs the fused CUDA launches and keeps the per-block operand buffers,
the backward runs the dgrad launches (esr_b200.engine.RRDBEngine.backward_input) and the exact adjoint of the CEM
projection (CEM_PyTorch.project_backward).  Weight gradients (training) are not built yet: if any generator
parameter requires grad this raises instead of silently falling back to an eager PyTorch graph."""
import torch


def _refuse_wgrad(net):
    if any(p.requires_grad for p in net.parameters()):
        raise NotImplementedError(
            'esr_b200: weight gradients (wgrad kernels) are not built yet. Freeze the generator '
            '(Z_optimizer.Manage_Model_Grad_Requirements / BaseModel.Set_Require_Grad_Status) or call under torch.no_grad(). '
            'There is deliberately no PyTorch/cuDNN fallback.')


class _RRDBFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net, pad):
        out, sv = net.engine().forward(x, pad=pad, save=True)
        ctx.net, ctx.sv = net, sv
        return out

    @staticmethod
    def backward(ctx, g_out):
        return ctx.net.engine().backward_input(g_out.contiguous(), ctx.sv), None, None


class _CemRRDBFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cem):
        net = cem.generated_image_model
        S = int(cem.ds_factor)
        pad = cem.invalidity_margins_LR if cem.pre_pad else 0
        G, sv = net.engine().forward(x, pad=pad, save=True)
        x_lr = x[:, -3:, :, :]
        if pad:
            x_lr = cem.LR_padder(x_lr)
        out = cem.project(x_lr, G, crop=pad * S)
        ctx.cem, ctx.net, ctx.sv, ctx.crop, ctx.hr_full = cem, net, sv, pad * S, (G.shape[2], G.shape[3])
        return out

    @staticmethod
    def backward(ctx, g_out):
        g_G, g_xlr = ctx.cem.project_backward(g_out.contiguous(), ctx.hr_full, crop=ctx.crop)
        gx = ctx.net.engine().backward_input(g_G, ctx.sv)
        if ctx.crop == 0:  # direct dependence of the projection on the LR image
            gx[:, -3:] += g_xlr
        return gx, None


def rrdb_forward_with_grad(net, x, pad):
    _refuse_wgrad(net)
    return _RRDBFn.apply(x, net, pad)


def cem_generator_forward_with_grad(cem, x):
    _refuse_wgrad(cem.generated_image_model)
    return _CemRRDBFn.apply(x, cem)
