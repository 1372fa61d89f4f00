"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes (the N>1 path of esr_b200.parallel)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'explorable-super-resolution_b200'))
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from esr_b200 import parallel
    torch.manual_seed(0)                                  # replicated weights
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.Conv2d(8, 3, 3, padding=1))
    frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)
    x = torch.randn(6, 3, 8, 8, generator=torch.Generator().manual_seed(1))   # the global batch, identical on all ranks
    a, b = parallel.shard_batch(6)
    # every rank: mean loss over ITS shard; averaging the gradients == gradient of the global-batch mean (equal shards)
    net(x[a:b]).abs().mean().backward()
    calls = parallel.average_gradients(list(net.parameters()) + [frozen], bucket_bytes=600)
    got = [p.grad.clone() for p in net.parameters()]
    ref_net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.Conv2d(8, 3, 3, padding=1))
    ref_net.load_state_dict(net.state_dict())
    ref_net(x).abs().mean().backward()
    err = max(float((g - p.grad).abs().max()) for g, p in zip(got, ref_net.parameters()))
    # relativistic-style global mean with unequal shards
    vals = torch.arange(7.0)
    s, e = parallel.shard_batch(7)
    gm = float(parallel.global_mean(vals[s:e]))
    # relativistic critic loss (SRRaGAN_model.py:353-354): means over the GLOBAL batch, differentiable across ranks
    import torch.nn.functional as F
    torch.manual_seed(5)
    critic = torch.nn.Linear(12, 1)
    real, fake = torch.randn(6, 12, generator=torch.Generator().manual_seed(2)), torch.randn(6, 12, generator=torch.Generator().manual_seed(3))

    def rel_loss(pr, pf, mean):
        return (F.binary_cross_entropy_with_logits(pr - mean(pf), torch.ones_like(pr)) +
                F.binary_cross_entropy_with_logits(pf - mean(pr), torch.zeros_like(pf))) / 2
    rel_loss(critic(real[a:b]), critic(fake[a:b]), parallel.global_mean_autograd).backward()
    parallel.average_gradients(list(critic.parameters()))
    got_c = [p.grad.clone() for p in critic.parameters()]
    ref_c = torch.nn.Linear(12, 1)
    ref_c.load_state_dict(critic.state_dict())
    rel_loss(ref_c(real), ref_c(fake), torch.mean).backward()
    err_rel = max(float((g - p.grad).abs().max()) for g, p in zip(got_c, ref_c.parameters()))
    # control-flow statistics are identical on every rank; replicas start from rank 0's weights
    stats = parallel.all_mean_scalars([float(rank), 2.0])
    gathered = parallel.all_gather_cat(torch.full((2,), float(rank)))
    lin = torch.nn.Linear(3, 2)
    torch.nn.init.constant_(lin.weight, float(rank + 1))
    parallel.broadcast_parameters(lin)
    ok_all, ok_one = parallel.agree(True), parallel.agree(rank == 0)
    assert stats == [0.5, 2.0] and gathered.tolist() == [0.0, 0.0, 1.0, 1.0] and float(lin.weight[0, 0]) == 1.0
    assert ok_all is True and ok_one is False
    q.put((rank, max(err, err_rel), calls, gm, (a, b), (s, e)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_averaging_and_global_mean_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for rank, err, calls, gm, shard, shard7 in out:
        assert err < 1e-6, err
        assert calls >= 2                      # several buckets were exercised
        assert abs(gm - 3.0) < 1e-6            # mean of 0..6 over ranks holding 4 and 3 values
    assert out[0][4] == (0, 3) and out[1][4] == (3, 6)
    assert out[0][5] == (0, 4) and out[1][5] == (4, 7)


def test_single_process_is_a_no_op():
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'explorable-super-resolution_b200'))
    from esr_b200 import parallel
    p = torch.nn.Parameter(torch.ones(4))
    p.grad = torch.full((4,), 2.0)
    assert parallel.average_gradients([p]) == 0 and float(p.grad[0]) == 2.0
    assert parallel.shard_batch(10, 4, 1) == (3, 6)
    assert float(parallel.global_mean(torch.tensor([1.0, 3.0]))) == 2.0
