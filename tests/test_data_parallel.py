"""Host-side logic of the N>1 path on CPU with the gloo backend (world_size 2): sharding, the flat gradient bucket and
the equivalence 'mean over ranks of per-shard gradients == gradient of the global batch'.  The model here is a plain
torch module -- the CUDA layers cannot run on CPU; what is under test is the data-parallel plumbing around them."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_indices_partition():
    from csmpn_b200.train_step import shard_indices

    for n, w in ((100, 8), (10, 4), (7, 2), (0, 2), (5, 8)):
        shards = [shard_indices(n, r, w) for r in range(w)]
        assert len({len(s) for s in shards}) == 1
        flat = sorted(i for s in shards for i in s)
        assert set(flat) == set(range(n))
        if n % w == 0:
            assert flat == list(range(n))
        dl = [shard_indices(n, r, w, drop_last=True) for r in range(w)]
        assert sum(len(s) for s in dl) == (n // w) * w and len({i for s in dl for i in s}) == (n // w) * w
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def test_flat_bucket_views():
    from csmpn_b200.train_step import FlatGradBucket

    m = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 2))
    b = FlatGradBucket(m.parameters())
    assert b.numel == sum(p.numel() for p in m.parameters())
    m(torch.randn(5, 3)).sum().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    assert torch.equal(ref, b.flat) and float(b.flat.abs().sum()) > 0
    assert all(p.grad.data_ptr() >= b.flat.data_ptr() for p in m.parameters())
    b.zero()
    assert all(float(p.grad.abs().sum()) == 0 for p in m.parameters())


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 1))

    def forward(self, batch, step, mode):
        per_sample = (self.net(batch["x"]).squeeze(-1) - batch["y"]) ** 2
        return per_sample.mean(), {"loss": per_sample}


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from csmpn_b200.train_step import DataParallelStep, shard_indices

        torch.manual_seed(0)
        model = _Toy()
        gen = torch.Generator().manual_seed(1)
        x, y = torch.randn(12, 6, generator=gen), torch.randn(12, generator=gen)
        ids = shard_indices(12, rank, world)
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        step = DataParallelStep(model, opt)
        before = [p.detach().clone() for p in model.parameters()]
        step({"x": x[ids], "y": y[ids]})
        assert all(p.grad.data_ptr() >= step.bucket.flat.data_ptr() for p in model.parameters())  # packed + attached
        ret[rank] = dict(grad=step.bucket.flat.clone(), params=[p.detach().clone() for p in model.parameters()], before=before)
    finally:
        dist.destroy_process_group()


def test_two_rank_step_equals_global_batch_step():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert torch.equal(ret[0]["grad"], ret[1]["grad"])
    torch.manual_seed(0)
    model = _Toy()
    gen = torch.Generator().manual_seed(1)
    x, y = torch.randn(12, 6, generator=gen), torch.randn(12, generator=gen)
    loss, _ = model({"x": x, "y": y}, 0, "train")
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(ret[0]["grad"], ref, rtol=1e-5, atol=1e-7)
    for p, q, b in zip(ret[0]["params"], ret[1]["params"], ret[0]["before"]):
        assert torch.equal(p, q) and not torch.equal(p, b)


def test_bucket_release_gather_roundtrip():
    """release() lets backward hand over its gradient tensors; gather() packs them (zeros for unused parameters)"""
    from csmpn_b200.train_step import FlatGradBucket

    m = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 2))
    unused = torch.nn.Parameter(torch.ones(5))
    params = list(m.parameters()) + [unused]
    b = FlatGradBucket(params)
    b.flat.fill_(7.0)
    b.release()
    assert all(p.grad is None for p in params)
    m(torch.randn(5, 3)).sum().backward()
    ref = [None if p.grad is None else p.grad.clone() for p in params]
    b.gather()
    off = 0
    for p, r in zip(params, ref):
        seg = b.flat[off: off + p.numel()].view_as(p)
        assert torch.equal(seg, torch.zeros_like(p) if r is None else r)
        assert p.grad.data_ptr() == seg.data_ptr()
        off += p.numel()
