"""Decoding strategies with the surface of rl4co/utils/decoding.py:191-461, backed by the
fused selection kernel (`co_select_action` = process_logits + Greedy/Sampling/Evaluate + logp
gather in one launch, no host synchronisation).

The reference performs >= 2 host syncs per step here (greedy/sampling asserts,
decoding.py:391-395,403-411); those asserts are dropped from the timed path -- infeasible
selections are impossible by construction because masked entries have probability 0.
"""

from __future__ import annotations

import torch

from . import native
from .ops import batchify, unbatchify, unbatchify_and_gather
from .tensordict import TensorDict


def get_log_likelihood(logprobs, actions=None, mask=None, return_sum: bool = True):
    """rl4co/utils/decoding.py:38-62 (the `> -1000` assert costs a sync and is omitted)."""
    if actions is not None and logprobs.dim() == 3:
        logprobs = logprobs.gather(-1, actions.unsqueeze(-1)).squeeze(-1)
    if mask is not None:
        logprobs = logprobs.masked_fill(~mask, 0)
    return logprobs.sum(1) if return_sum else logprobs


def keep_top_k(logits: torch.Tensor, top_k: int) -> torch.Tensor:
    """rl4co/utils/decoding.py:109-114 -- logits below the row's k-th largest become -inf (out of place)."""
    kth = torch.topk(logits, top_k).values[..., -1:]
    return logits.masked_fill(logits < kth, float("-inf"))


def keep_top_p(logits: torch.Tensor, top_p: float) -> torch.Tensor:
    """rl4co/utils/decoding.py:117-135 -- nucleus filtering with the reference's exact operation order (ascending
    sort -> softmax -> cumsum -> drop while the cumulative mass is <= 1 - top_p), so the kept sets are identical."""
    if not 0.0 < top_p < 1.0:
        return logits
    ordered, order = torch.sort(logits, dim=-1)
    tail = ordered.softmax(dim=-1).cumsum(dim=-1) <= (1.0 - top_p)
    return logits.masked_fill(torch.zeros_like(tail).scatter(-1, order, tail), float("-inf"))


class DecodingStrategy:
    """rl4co/utils/decoding.py:191-423"""

    name = "base"
    select_mode = native.SELECT_GREEDY

    def __init__(self, temperature: float = 1.0, top_p: float = 0.0, top_k: int = 0, mask_logits: bool = True,
                 tanh_clipping: float = 0, num_samples: int | None = None, multisample: bool = False,
                 num_starts: int | None = None, multistart: bool = False, select_start_nodes_fn=None,
                 improvement_method_mode: bool = False, select_best: bool = False, store_all_logp: bool = False,
                 **kwargs) -> None:
        assert top_p <= 1.0, "top-p should be in (0, 1]."
        self.top_p, self.top_k = top_p, top_k
        if improvement_method_mode:
            raise NotImplementedError("improvement_method_mode is outside the fused path")
        self.temperature, self.mask_logits, self.tanh_clipping = temperature, mask_logits, tanh_clipping
        assert not (multistart and multisample), "Using both multistart and multisample is not supported"
        if num_samples and num_starts:
            assert not (num_samples > 1 and num_starts > 1)
        if num_samples is not None:
            multisample = num_samples > 1
        if num_starts is not None:
            multistart = num_starts > 1
        self.multistart, self.multisample = multistart, multisample
        self.num_starts = num_starts if multistart else num_samples
        self.select_start_nodes_fn = select_start_nodes_fn
        self.select_best = select_best
        self.store_all_logp = store_all_logp
        self.actions, self.logprobs = [], []

    def pre_decoder_hook(self, td: TensorDict, env, action: torch.Tensor | None = None):
        """rl4co/utils/decoding.py:282-330"""
        if self.multistart or self.multisample:
            if self.num_starts is None:
                self.num_starts = env.get_num_starts(td)
        else:
            self.num_starts = 0
        if self.num_starts >= 1:
            if self.multistart:
                if action is None:
                    if self.select_start_nodes_fn is not None:
                        action = self.select_start_nodes_fn(td, env, self.num_starts)
                    else:
                        action = env.select_start_nodes(td, num_starts=self.num_starts)
                td = batchify(td, self.num_starts)
                td.set("action", action)
                td = env.step(td)["next"]
                if self.store_all_logp:
                    logprobs = torch.zeros_like(td["action_mask"], dtype=torch.float32)
                else:
                    logprobs = torch.zeros_like(action, dtype=torch.float32)
                self.logprobs.append(logprobs)
                self.actions.append(action)
            else:
                td = batchify(td, self.num_starts)
        return td, env, self.num_starts

    def post_decoder_hook(self, td: TensorDict, env):
        """rl4co/utils/decoding.py:332-342"""
        assert len(self.logprobs) > 0, "No logprobs were collected because all environments were done"
        logprobs = torch.stack(self.logprobs, 1)
        actions = torch.stack(self.actions, 1)
        if self.num_starts > 0 and self.select_best:
            logprobs, actions, td, env = self._select_best(logprobs, actions, td, env)
        return logprobs, actions, td, env

    def _noise(self, logits):
        return None

    def step(self, logits: torch.Tensor, mask: torch.Tensor, td: TensorDict | None = None,
             action: torch.Tensor | None = None, **kwargs) -> TensorDict:
        """rl4co/utils/decoding.py:344-385, one kernel."""
        assert td is not None, "td must be provided"
        act_io = action.contiguous().clone() if self.select_mode == native.SELECT_EVALUATE else None
        tanh_clipping, temperature, mask_logits = self.tanh_clipping, self.temperature, self.mask_logits
        if self.top_k > 0 or self.top_p > 0:
            # process_logits (decoding.py:168-185) up to the filters with library ops (top-k / sort are not worth
            # a kernel at N <= 128); the kept set then goes to the selection kernel as its mask, which finishes
            # with log_softmax over exactly those entries
            z = torch.tanh(logits) * tanh_clipping if tanh_clipping > 0 else logits
            if mask_logits:
                assert mask is not None, "mask must be provided if mask_logits is True"
                z = z.masked_fill(~mask, float("-inf"))
            z = z / temperature
            if self.top_k > 0:
                z = keep_top_k(z, min(self.top_k, z.size(-1)))
            if self.top_p > 0:
                z = keep_top_p(z, self.top_p)
            logits, mask = z, torch.isfinite(z)
            tanh_clipping, temperature, mask_logits = 0.0, 1.0, True
        selected, logp, all_lp = native.select_action(
            logits.contiguous(), mask.contiguous() if mask is not None else None, self.select_mode,
            noise=self._noise(logits), action=act_io, tanh_clipping=tanh_clipping, temperature=temperature,
            mask_logits=mask_logits, store_all_logp=self.store_all_logp)
        td.set("action", selected)
        self.actions.append(selected)
        self.logprobs.append(all_lp if self.store_all_logp else logp)
        return td

    def _select_best(self, logprobs, actions, td: TensorDict, env):
        """rl4co/utils/decoding.py:415-423"""
        rewards = env.get_reward(td, actions)
        _, max_idxs = unbatchify(rewards, self.num_starts).max(dim=-1)
        actions = unbatchify_and_gather(actions, max_idxs, self.num_starts)
        logprobs = unbatchify_and_gather(logprobs, max_idxs, self.num_starts)
        td = unbatchify_and_gather(td, max_idxs, self.num_starts)
        return logprobs, actions, td, env


class Greedy(DecodingStrategy):
    name = "greedy"
    select_mode = native.SELECT_GREEDY


class Sampling(DecodingStrategy):
    """torch.multinomial(p, 1) draws one `empty_like(p).exponential_(1)` and returns
    argmax(p / q); drawing q from the same torch CUDA generator and doing the division +
    arg-max in-kernel consumes the generator identically."""

    name = "sampling"
    select_mode = native.SELECT_SAMPLE_NOISE

    def _noise(self, logits):
        return torch.empty_like(logits).exponential_(1)


class Evaluate(DecodingStrategy):
    name = "evaluate"
    select_mode = native.SELECT_EVALUATE


def get_decoding_strategy(decoding_strategy, **config) -> DecodingStrategy:
    """rl4co/utils/decoding.py:17-35 (beam search is outside the fused path)."""
    registry = {"greedy": Greedy, "sampling": Sampling, "multistart_greedy": Greedy,
                "multistart_sampling": Sampling, "evaluate": Evaluate}
    if decoding_strategy not in registry:
        raise NotImplementedError(f"decode type {decoding_strategy!r} is outside the fused path: {list(registry)}")
    if "multistart" in decoding_strategy:
        config["multistart"] = True
    return registry[decoding_strategy](**config)
