"""Gradient path of the generator (+CEM): input gradient (what the latent-exploration loop needs,
Z_optimization.py:673-749: loss.backward() through a frozen G down to the latent map Z) and weight gradients
(training, models/SRRaGAN_model.py:436-500: l_g_total.backward(); optimizer_G.step()).

torch.autograd only sees one node: the forward runs the fused CUDA launches and keeps the per-block operand buffers,
the backward runs the dgrad (+ wgrad) launches (esr_b200.engine.RRDBEngine.backward) and the exact adjoint of the CEM
projection (CEM_PyTorch.project_backward).  The generator's parameters are inputs of that node, so .grad accumulation,
optimizers, gradient hooks and DistributedDataParallel behave exactly as with the reference's nn.Conv2d graph.  There is
no PyTorch/cuDNN fallback."""
import torch


def _params(net):
    """(weight, bias) of every conv in engine order, flattened"""
    out = []
    for c in net.engine()._convs():
        out += [c.weight, c.bias]
    return out


def _engine_for(net, params):
    """training (some generator parameter wants a gradient) runs in bf16, the frozen generator in net.compute_dtype;
    the parity mode of esr_b200.precision runs both in split precision"""
    training = any(p.requires_grad for p in params)
    return net.engine(training=training)


def _flat_grads(ctx, grads, n_params):
    """engine grads [(dW, db)] -> one entry per parameter input (None where autograd does not need it)"""
    flat = []
    for k in range(n_params):
        need = ctx.needs_input_grad[ctx.first_param + k]
        g = grads[k // 2][k % 2] if (grads is not None and need) else None
        if isinstance(g, str):      # engine.DIRECT: already written into the parameter's registered flat-buffer view (.grad points at it)
            g = None
        flat.append(g)
    return flat


class _RRDBFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net, pad, *params):
        eng = _engine_for(net, params)
        out, sv = eng.forward(x, pad=pad, save=True)
        ctx.eng, ctx.sv, ctx.first_param, ctx.n_params = eng, sv, 3, len(params)
        return out

    @staticmethod
    def backward(ctx, g_out):
        wgrad = any(ctx.needs_input_grad[ctx.first_param:])
        gx, grads = ctx.eng.backward(g_out.contiguous(), ctx.sv, wgrad=wgrad)
        return (gx if ctx.needs_input_grad[0] else None, None, None) + tuple(_flat_grads(ctx, grads, ctx.n_params))


class _CemRRDBFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cem, *params):
        net = cem.generated_image_model
        S = int(cem.ds_factor)
        pad = cem.invalidity_margins_LR if cem.pre_pad else 0
        eng = _engine_for(net, params)
        G, sv = eng.forward(x, pad=pad, save=True)
        x_lr = x[:, -3:, :, :]
        if pad:
            x_lr = cem.LR_padder(x_lr)
        out = cem.project(x_lr, G, crop=pad * S)
        ctx.cem, ctx.eng, ctx.sv, ctx.crop, ctx.hr_full = cem, eng, sv, pad * S, (G.shape[2], G.shape[3])
        ctx.first_param, ctx.n_params = 2, len(params)
        return out

    @staticmethod
    def backward(ctx, g_out):
        g_G, g_xlr = ctx.cem.project_backward(g_out.contiguous(), ctx.hr_full, crop=ctx.crop)
        wgrad = any(ctx.needs_input_grad[ctx.first_param:])
        gx, grads = ctx.eng.backward(g_G, ctx.sv, wgrad=wgrad)
        if ctx.crop == 0:  # direct dependence of the projection on the LR image
            gx[:, -3:] += g_xlr
        return (gx if ctx.needs_input_grad[0] else None, None) + tuple(_flat_grads(ctx, grads, ctx.n_params))


def rrdb_forward_with_grad(net, x, pad):
    return _RRDBFn.apply(x, net, pad, *_params(net))


def cem_generator_forward_with_grad(cem, x):
    return _CemRRDBFn.apply(x, cem, *_params(cem.generated_image_model))
