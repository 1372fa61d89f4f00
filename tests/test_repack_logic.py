"""Host logic of the in-place weight re-packing (esr_b200.engine.RRDBEngine.packed / packed_t / packed_bwd_dense): the packed
objects are built once, refreshed in place (one queue flush per group) after the parameters change, and keep their identity so
that recorded launch plans stay valid.  The CUDA entry points are replaced by recorders."""
import torch


class _FakePacked:
    built = 0

    def __init__(self, weight, bias, dtype=None, lead=0, transpose_flip=False, queue=None, **kw):
        _FakePacked.built += 1
        self.wpacked = weight.detach()
        self.shape, self.t, self.repacks = tuple(weight.shape), transpose_flip, 0
        self.last = weight.detach().clone()
        queue.append(self)

    def repack(self, weight, bias, queue=None):
        assert tuple(weight.shape) == self.shape
        self.repacks += 1
        self.last = weight.detach().clone()
        queue.append(self)


def test_packed_objects_persist_and_refresh_in_place(monkeypatch):
    import models.modules.architecture as arch
    from esr_b200 import ops
    flushes = []
    monkeypatch.setattr(ops, 'PackedConv', _FakePacked)
    monkeypatch.setattr(ops, 'run_pack_queue', lambda q: (flushes.append(len(q)), q.clear()))
    net = arch.RRDBNet(3, 3, 32, 2, upscale=4, num_latent_channels=0)
    eng = net.engine(torch.bfloat16)
    n_conv = len(eng._convs())
    assert n_conv == 1 + 2 * 15 + 1 + 2 + 2
    _FakePacked.built = 0
    pk = eng.packed()
    pt = eng.packed_t()
    bd = eng.packed_bwd_dense()
    assert len(pk) == n_conv and sum(p is not None for p in pt) == n_conv - 30 and len(bd) == 6 and all(len(r) == 5 for r in bd)
    assert _FakePacked.built == n_conv + (n_conv - 30) + 30 and flushes == [n_conv, n_conv - 30, 30]
    # nothing changed: nothing is re-packed, same objects
    assert eng.packed() is pk and eng.packed_t() is pt and eng.packed_bwd_dense() is bd
    assert [f for f in flushes if f] == [n_conv, n_conv - 30, 30]
    before = list(flushes)
    # an optimizer step: every group refreshed in place by ONE flush each, identities kept
    with torch.no_grad():
        for p in net.parameters():
            p.add_(1.0)
    pk2, pt2, bd2 = eng.packed(), eng.packed_t(), eng.packed_bwd_dense()
    assert pk2 is pk and pt2 is pt
    assert all(a is b for ra, rb in zip(bd, bd2) for a, b in zip(ra, rb))
    assert _FakePacked.built == n_conv + (n_conv - 30) + 30            # no new objects
    assert all(p.repacks == 1 for p in pk) and all(p.repacks == 1 for p in pt if p is not None) and all(p.repacks == 1 for r in bd for p in r)
    assert [f for f in flushes[len(before):] if f] == [n_conv, n_conv - 30, 30]
    assert torch.equal(pk[0].last, net.model[0].weight.detach())
    # the refreshed dense-backward weights are the new combination
    from esr_b200.engine import combine_dense_backward_weights
    convs = eng._convs()
    W = [torch.stack([convs[1 + r * 5 + jj].weight.detach().float() for r in range(6)]) for jj in range(5)]
    a5 = torch.tensor([0.04 if r % 3 == 2 else 0.2 for r in range(6)])
    comb = combine_dense_backward_weights(W, a5, 0, 32, 32)
    assert torch.equal(bd[3][2].last, comb[2][3])
