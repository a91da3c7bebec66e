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
