"""Multi-GPU plumbing of the rollout path.

The path shards by instance: rank r owns a contiguous slice of the batch (all augmentations
and all starts of an instance stay on one GPU), weights are replicated, and nothing crosses
GPUs while episodes run.  The only exchange is REINFORCE's mean baseline: one NCCL all-reduce
of {sum(reward), count} in float64 per training step, so that the baseline equals the
single-process reference on the concatenated batch (rl4co/models/rl/reinforce/baselines.py:
75-81 takes `reward.mean()`; under the reference's DDP that mean is per rank -- a deliberate,
documented difference).
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split of `total` instances; the first `total % world` ranks get one more."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_tensordict(td, rank: int | None = None, world: int | None = None):
    """Slice a TensorDict of instances for this rank (before aug / multistart expansion)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(td.batch_size[0], rank, world)
    return td[lo:hi]


def local_reward_stats(reward: torch.Tensor) -> torch.Tensor:
    """{sum, count} as float64[2].  CUDA tensors go through co_reward_stats (one kernel);
    this function is also the unit under the gloo test, where the statistics of a host tensor
    are formed by the test itself and passed to `allreduce_mean`."""
    from . import native

    out = torch.zeros(2, dtype=torch.float64, device=reward.device)
    native.reward_stats(reward.contiguous(), out)
    return out


def allreduce_mean(stats: torch.Tensor) -> torch.Tensor:
    """Sum {sum, count} over ranks (NCCL on GPU, gloo in the CPU tests) -> global mean (f32)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return (stats[0] / stats[1]).to(torch.float32)


def global_mean_baseline(reward: torch.Tensor) -> torch.Tensor:
    return allreduce_mean(local_reward_stats(reward))


def sync_gradients(parameters) -> int:
    """Average the gradients of `parameters` over ranks (what DDP's reducer does for a wrapped module,
    rl4co/utils/trainer.py:83-86).  `reinforce_step` / `pomo_step` drive the encoder / decoder sub-modules
    directly, so a DDP wrapper's forward hooks never fire -- the training steps call this before gradient
    clipping and `optimizer.step()` instead.  One flat all-reduce (2.8 MB for the 0.71 M-parameter AM), then
    scatter back.  Parameters without a gradient contribute zeros so every rank reduces the same layout.
    Returns the number of elements reduced (0 when not distributed)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return 0
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return off


def rank_stream_offset() -> int:
    """Philox stream selector of this rank (word 3 of the counter, csrc/co_common.cuh:philox_exp1): ranks that
    were seeded identically still draw independent noise streams."""
    return dist.get_rank() if (dist.is_initialized() and dist.get_world_size() > 1) else 0
