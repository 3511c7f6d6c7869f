"""Host-side multi-rank logic on CPU with the gloo backend (world_size 2): ray sharding + the single flat-buffer
gradient all-reduce.  The kernels themselves need a GPU; what is checked here is that k ranks working on disjoint
ray slices and summing one flat buffer reproduce the single-rank gradient of a shard-linear loss."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cope_nerf_b200.dist import FlatGradBucket, allreduce_scalar_, shard_range


def test_shard_range_alignment_and_cover():
    for n, world in ((1024, 1), (1024, 2), (1024, 8), (1000, 3), (16, 8), (4096 + 7, 4)):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
            assert b == c
        for a, b in spans[:-1]:
            assert a % 16 == 0 and (b - a) % 16 == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Softplus(beta=100), torch.nn.Linear(7, 3))
    x, y = torch.randn(64, 5), torch.randn(64, 3)
    bucket = FlatGradBucket(net.parameters())
    a, b = shard_range(64, rank, world)
    # shard-linear loss (sum / global N), like rgb L1 and the eikonal mean
    loss = (net(x[a:b]) - y[a:b]).abs().sum() / 64.0
    bucket.zero_()
    loss.backward()
    bucket.allreduce_()
    w = allreduce_scalar_(torch.tensor([float(b - a)]))
    q.put((rank, bucket.flat.clone(), w.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_allreduce_equals_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Softplus(beta=100), torch.nn.Linear(7, 3))
    x, y = torch.randn(64, 5), torch.randn(64, 3)
    ((net(x) - y).abs().sum() / 64.0).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    for rank, flat, w in got:
        assert torch.allclose(flat, ref, atol=1e-6), rank
        assert w == 64.0


def test_flat_grad_bucket_keeps_and_reattaches_gradients():
    """ADVICE (round 1): building the bucket must not discard gradients that already exist, and gradients that autograd
    allocated outside the bucket (optimizer.zero_grad(set_to_none=True) dropped the views) must be pulled back in by
    allreduce_() instead of a stale buffer being exchanged."""
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.Tanh(), torch.nn.Linear(6, 2))
    x = torch.randn(10, 4)
    net(x).pow(2).sum().backward()
    ref = torch.cat([p.grad.reshape(-1).clone() for p in net.parameters()])
    bucket = FlatGradBucket(net.parameters())                  # after backward(): the gradients are copied in, not zeroed
    assert torch.allclose(bucket.flat, ref)
    for p in net.parameters():
        assert p.grad.data_ptr() >= bucket.flat.data_ptr() and p.grad.data_ptr() < bucket.flat.data_ptr() + bucket.flat.numel() * 4
    opt = torch.optim.SGD(net.parameters(), lr=0.0)
    opt.zero_grad()                                            # set_to_none=True: the views are gone
    assert all(p.grad is None for p in net.parameters())
    net(x).pow(2).sum().backward()                             # fresh .grad tensors outside the bucket
    bucket.allreduce_()                                        # single rank: no exchange, but the re-attach must happen
    assert torch.allclose(bucket.flat, ref)
    bucket.zero_()
    assert float(bucket.flat.abs().max()) == 0.0 and all(float(p.grad.abs().max()) == 0.0 for p in net.parameters())
    net(x).pow(2).sum().backward()                             # accumulates into the views again
    assert torch.allclose(bucket.flat, ref)


def test_flatten_params_and_param_range():
    """FlatGradBucket.flatten_params_: every p.data becomes a view of ONE flat buffer laid out like the gradients (values kept,
    forward unchanged, idempotent); param_range: contiguous runs only.  optim.FlatAdam has no CPU path."""
    import cope_nerf_b200 as C
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    before = [p.detach().clone() for p in net.parameters()]
    x = torch.randn(4, 5)
    y0 = net(x).detach().clone()
    bucket = FlatGradBucket(list(net.parameters()))
    fp = bucket.flatten_params_()
    assert bucket.flatten_params_() is fp and fp.numel() == bucket.flat.numel() == 66
    off = 0
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p.detach(), b) and p.data_ptr() == fp.data_ptr() + 4 * off
        off += p.numel()
    assert torch.equal(net(x), y0)
    net(x).sum().backward()                                  # gradients still land in the bucket's views
    assert bucket.flat.abs().sum() > 0
    fp.mul_(2.0)                                             # an update of the flat buffer IS an update of the module
    assert torch.allclose(net[0].weight.detach(), before[0] * 2)
    assert bucket.param_range(list(net[1].parameters())) == (42, 66)
    assert bucket.param_range(list(net.parameters())) == (0, 66)
    with pytest.raises(ValueError):
        bucket.param_range([net[0].weight, net[1].weight])   # not a contiguous run
    assert bucket.params_attached()
    opt = C.optim.FlatAdam(bucket, lr=1e-3, params=list(net[1].parameters()))
    assert opt.range == (42, 66) and isinstance(opt, torch.optim.Optimizer)
    torch.optim.lr_scheduler.MultiStepLR(opt, [1], 0.1)      # schedulers attach (train.py:61-64)
    with pytest.raises(C.CopeError):
        opt.step()                                           # CPU tensors: no fallback
    net[0].weight.data = net[0].weight.data.clone()          # what module.half().float() / .to(device) do to the storage
    assert not bucket.params_attached()                      # ... FlatAdam.step refuses instead of updating a stale buffer
