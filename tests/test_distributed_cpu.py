"""world_size-2 gloo tests (CPU) of the N>1 host logic: instance sharding and the baseline
{sum,count} all-reduce reproduce the single-process mean on the concatenated batch."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rl4co_b200.distributed import allreduce_mean, shard_bounds, shard_tensordict
from rl4co_b200.tensordict import TensorDict


def test_shard_bounds_cover_and_are_contiguous():
    for total in (1, 7, 8, 65536, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    reward = -torch.rand(total) * 10  # identical on both ranks: the "global" batch
    td = TensorDict({"locs": torch.arange(total * 4, dtype=torch.float32).view(total, 2, 2), "r": reward}, batch_size=[total])
    mine = shard_tensordict(td)
    lo, hi = shard_bounds(total, rank, world)
    assert torch.equal(mine["locs"], td["locs"][lo:hi])
    r = mine["r"]
    stats = torch.tensor([r.double().sum().item(), float(r.numel())], dtype=torch.float64)
    mean = allreduce_mean(stats)
    q.put((rank, mean.item(), reward.mean().item(), hi - lo))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_baseline_allreduce_equals_single_process_mean():
    world, total = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert sum(x[3] for x in res) == total
    for _, mean, ref, _ in res:
        assert abs(mean - ref) < 1e-6
    assert res[0][1] == res[1][1]  # identical on both ranks


def _grad_worker(rank, world, port, q):
    import torch.nn as nn

    from rl4co_b200.distributed import rank_stream_offset, sync_gradients

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # identical replicas
    net = nn.Sequential(nn.Linear(4, 8), nn.ReLU(), nn.Linear(8, 1))
    frozen = nn.Linear(3, 3)  # a parameter that gets no gradient on any rank
    opt = torch.optim.SGD(list(net.parameters()) + list(frozen.parameters()), lr=0.1)
    torch.manual_seed(100 + rank)  # different data per rank
    x = torch.randn(16, 4)
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        net(x).pow(2).mean().backward()
        n = sync_gradients(list(net.parameters()) + list(frozen.parameters()))
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    q.put((rank, flat.tolist(), n, rank_stream_offset()))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_sync_keeps_replicas_identical():
    """reinforce_step / pomo_step bypass DDP, so they average gradients themselves; after a few
    steps on different data the replicas' parameters must still be bit-identical."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]
    assert res[0][2] == res[1][2] > 0
    assert [r[3] for r in res] == [0, 1]  # distinct Philox streams per rank
