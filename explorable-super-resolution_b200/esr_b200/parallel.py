"""Multi-GPU plumbing: one process per GPU (torchrun), batch sharded by rank, replicated weights.

The forward / test / Z-optimisation paths shard by sample with no data-path collective.  Training adds
  * one all-reduce of the generator's gradients per optimizer step (`average_gradients`: flat fp32 buckets over NCCL /
    NVLink; the generator's parameters are ordinary autograd inputs of the fused node, so DistributedDataParallel works
    as well), and
  * a 2-scalar all-reduce where the reference takes a mean over the GLOBAL batch after nn.DataParallel gathered the
    discriminator outputs on GPU 0 (models/SRRaGAN_model.py:353-354,475-476): `global_mean`.
Everything here is backend-agnostic torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def rank():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def shard_batch(n, world_size=None, r=None):
    """[start, stop) of the samples of a global batch of n owned by rank r (contiguous, sizes differ by at most 1)"""
    world_size = world() if world_size is None else world_size
    r = rank() if r is None else r
    base, rem = divmod(n, world_size)
    start = r * base + min(r, rem)
    return start, start + base + (1 if r < rem else 0)


def average_gradients(params, bucket_bytes=64 << 20, group=None, optimizer=None):
    """In-place average of .grad over all ranks, in flat buckets of about `bucket_bytes` (a 17 M-parameter generator is two
    buckets: the all-reduce is latency-bound on NVLink).  Parameters without a gradient contribute zeros so
    that every rank issues the same collectives.  Returns the number of all-reduce calls.
    With a FlatAdam `optimizer` whose gradients already live in its flat buffer (the engines write them there) the buffer itself is
    reduced in place: no concatenation, no copy back."""
    params = [p for p in params if p.requires_grad]
    if world() == 1 or not params:
        return 0
    if optimizer is not None and hasattr(optimizer, 'flat_grad'):
        calls = 0
        flat_ok = True
        for gi in range(len(optimizer.param_groups)):
            flat_ok = flat_ok and optimizer.flat_grad(gi) is not None and optimizer.grads_in_place(gi)
        flat_ok = agree(flat_ok, group=group)      # every rank takes the same path
        if flat_ok:
            for gi in range(len(optimizer.param_groups)):
                buf = optimizer.flat_grad(gi)
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
                buf.div_(world())
                calls += 1
            return calls
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world())
        off = 0
        for p in bucket:
            n = p.numel()
            g = flat[off:off + n].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        calls += 1
        bucket, size = [], 0

    for p in params:
        bucket.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return calls


def broadcast_parameters(module, src=0, group=None):
    """every rank starts from rank `src`'s parameters and buffers (the reference's nn.DataParallel replicates GPU 0's weights
    every forward, models/networks.py:122; with one process per GPU the replicas only agree if they start equal)"""
    if world() == 1:
        return 0
    n = 0
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
            n += 1
    return n


def all_mean_scalars(values, group=None):
    """mean over the ranks of a list of host floats (the loss / logit statistics the reference computes on the gathered global
    batch, models/SRRaGAN_model.py:372-381; equal shard sizes make the mean of rank means the global mean).  Every statistic that
    gates control flow (D verification, lr roll-back) goes through here so that all ranks take the same branch."""
    if world() == 1 or len(values) == 0:
        return [float(v) for v in values]
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(v) / world() for v in t.cpu()]


def all_gather_cat(t, group=None):
    """per-sample values of every rank, concatenated in rank order (equal shard sizes)"""
    if world() == 1:
        return t
    out = [torch.empty_like(t) for _ in range(world())]
    dist.all_gather(out, t.contiguous(), group=group)
    return torch.cat(out, 0)


def agree(flag, group=None):
    """True only if `flag` is true on every rank (a safety net behind the shared statistics: ranks never split on a branch that
    contains collectives)"""
    if world() == 1:
        return bool(flag)
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.cpu()[0]))


def global_mean(local_values, group=None):
    """mean over the GLOBAL batch of per-sample values held rank by rank (unequal shard sizes allowed): one 2-scalar
    all-reduce (sum, count)"""
    s = torch.stack([local_values.sum().float(), torch.tensor(float(local_values.numel()), device=local_values.device)])
    if world() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    return s[0] / s[1]


class _GlobalMeanFn(torch.autograd.Function):
    """differentiable mean over the global batch.  Forward: one 2-scalar all-reduce.  Backward: every rank's loss depends on
    the mean, so the gradient reaching it is summed over the ranks (one scalar all-reduce) before it is spread over the local
    values; with rank-mean losses and averaged parameter gradients this reproduces the single-process gradient exactly."""

    @staticmethod
    def forward(ctx, t):
        s = torch.stack([t.sum().float(), torch.tensor(float(t.numel()), device=t.device)])
        if world() > 1:
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
        ctx.n, ctx.shape = float(s[1]), t.shape
        return s[0] / s[1]

    @staticmethod
    def backward(ctx, g):
        g = g.clone().float()
        if world() > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return (g / ctx.n).expand(ctx.shape)


def global_mean_autograd(local_values):
    return _GlobalMeanFn.apply(local_values)
