"""Tensor helpers of the rollout path, same names / argument meaning as rl4co/utils/ops.py.

Pure layout helpers (batchify / unbatchify / gather_by_index / start-node selection) are
views and index arithmetic in torch on whatever device the data lives; the arithmetic op
(`get_tour_length`-based rewards) goes through the CUDA library (rl4co_b200.native).
"""

from __future__ import annotations

import torch

from .tensordict import TensorDict


def _batchify_single(x, repeats: int):
    """rl4co/utils/ops.py:10-13: start-major repeat, flat index = r * B + b."""
    s = x.shape
    return x.expand(repeats, *s).contiguous().view(s[0] * repeats, *s[1:])


def batchify(x, shape):
    """rl4co/utils/ops.py:16-29"""
    shape = [shape] if isinstance(shape, int) else shape
    for s in reversed(shape):
        x = _batchify_single(x, s) if s > 0 else x
    return x


def _unbatchify_single(x, repeats: int):
    """rl4co/utils/ops.py:32-35"""
    s = x.shape
    return x.view(repeats, s[0] // repeats, *s[1:]).permute(1, 0, *range(2, len(s) + 1))


def unbatchify(x, shape):
    """rl4co/utils/ops.py:38-51: '(r b) ... -> b r ...'"""
    shape = [shape] if isinstance(shape, int) else shape
    for s in reversed(shape):
        x = _unbatchify_single(x, s) if s > 0 else x
    return x


def gather_by_index(src, idx, dim=1, squeeze=True):
    """rl4co/utils/ops.py:54-66"""
    expanded_shape = list(src.shape)
    expanded_shape[dim] = -1
    idx = idx.view(idx.shape + (1,) * (src.dim() - idx.dim())).expand(expanded_shape)
    squeeze = idx.size(dim) == 1 and squeeze
    return src.gather(dim, idx).squeeze(dim) if squeeze else src.gather(dim, idx)


def unbatchify_and_gather(x, idx, n: int):
    """rl4co/utils/ops.py:69-74"""
    x = unbatchify(x, n)
    return gather_by_index(x, idx, dim=idx.dim())


def get_num_starts(td, env_name=None) -> int:
    """rl4co/utils/ops.py:115-125 (tsp / cvrp / sdvrp / op rows)"""
    num_starts = td["action_mask"].shape[-1]
    if env_name in ("cvrp", "sdvrp", "op", "pctsp"):
        num_starts -= 1
    return num_starts


def select_start_nodes(td, env, num_starts: int):
    """rl4co/utils/ops.py:128-149 (tsp / depot-env rows)"""
    num_loc = env.generator.num_loc if hasattr(env.generator, "num_loc") else 0xFFFFFFFF
    sel = torch.arange(num_starts, device=td.device).repeat_interleave(td.shape[0]) % num_loc
    if env.name == "tsp":
        return sel
    sel = sel + 1
    if env.name == "op" and (td["action_mask"][..., 1:].float().sum(-1) < num_starts).any():
        # ops.py:150-160: some customers may be out of reach: resample the starts from the available ones
        sel = torch.multinomial(td["action_mask"][..., 1:].float(), num_starts, replacement=True) + 1
        sel = sel.t().reshape(-1)  # "b n -> (n b)"
    return sel


def get_tour_length_reward(locs, actions, with_depot: bool):
    """-get_tour_length(gather_by_index(locs, actions)) fused in one CUDA kernel
    (rl4co/utils/ops.py:54-90 as used by tsp/env.py:150-156 and cvrp/env.py:138-147)."""
    from . import native

    return native.tour_length(locs.contiguous(), actions.contiguous(), with_depot)


def dihedral_8_augmentation(xy):
    """rl4co/data/transforms.py:16-38 (aug-major: flat index = a * B + b)."""
    x, y = xy.split(1, dim=2)
    zs = ((x, y), (1 - x, y), (x, 1 - y), (1 - x, 1 - y), (y, x), (1 - y, x), (y, 1 - x), (1 - y, 1 - x))
    return torch.cat([torch.cat(z, dim=2) for z in zs], dim=0)


class StateAugmentation:
    """rl4co/data/transforms.py:113-151 restricted to the dihedral-8 function POMO uses."""

    def __init__(self, num_augment: int = 8, augment_fn: str = "dihedral8", feats=None, **_):
        assert augment_fn == "dihedral8" and num_augment == 8, "only dihedral8 x8 is on the hot path"
        self.num_augment = num_augment
        self.feats = ["locs"] if feats is None else feats

    def __call__(self, td: TensorDict) -> TensorDict:
        td_aug = batchify(td, self.num_augment)
        for feat in self.feats:
            x = td_aug[feat]
            base = x[: x.shape[0] // 8]
            if base.is_cuda and base.dim() == 3 and base.shape[-1] == 2 and base.dtype == torch.float32:
                from . import native

                td_aug.set(feat, native.dihedral8(base.contiguous()))  # one kernel: 8 B read, 64 B written per node
            else:
                td_aug.set(feat, dihedral_8_augmentation(base))
        return td_aug
