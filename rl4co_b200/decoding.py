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


def beam_expand(logprobs: torch.Tensor, cum: torch.Tensor, beam_width: int):
    """One beam step (rl4co/utils/decoding.py:568-600) on the flat beam-major batch (row = w * B + b).

    logprobs [B*W, N] full log-prob rows, cum [B*W, 1] cumulative log-prob of each live beam. Per instance the W best
    (parent beam, node) pairs are kept. Returns (node [B*W], parent [B*W] int32, src_row [B*W], new cum [B*W, 1])."""
    rows, n = logprobs.shape
    b = rows // beam_width
    # the candidates of instance b side by side: column w * N + node
    cand = (logprobs + cum).view(beam_width, b, n).permute(1, 0, 2).reshape(b, beam_width * n)
    best, flat = torch.topk(cand, beam_width, dim=1)
    best, flat = best.t().reshape(-1), flat.t().reshape(-1)  # back to beam-major rows
    parent = torch.div(flat, n, rounding_mode="floor").int()
    src = torch.arange(b, device=logprobs.device).repeat(beam_width) + parent * b
    return flat % n, parent, src, best.unsqueeze(1)


def beam_backtrack(actions: torch.Tensor, logprobs: torch.Tensor, parents: torch.Tensor, beam_width: int):
    """Re-thread the per-step records along the parent pointers, last step first (rl4co/utils/decoding.py:529-556).

    actions [B*W, T], logprobs [B*W, T, N] (or [B*W, T]), parents [B*W, T] int. Row r of the result is the complete
    sequence that ends in final beam r."""
    rows, T = actions.shape
    b = rows // beam_width
    inst = torch.arange(b, device=actions.device).repeat(beam_width)
    out_a, out_lp = torch.empty_like(actions), torch.empty_like(logprobs)
    src = torch.arange(rows, device=actions.device)  # the final beams read their own last record
    for k in range(T - 1, -1, -1):
        out_a[:, k], out_lp[:, k] = actions[src, k], logprobs[src, k]
        if k:
            src = inst + parents[src, k].long() * b
    return out_a, out_lp


class BeamSearch(DecodingStrategy):
    """rl4co/utils/decoding.py:464-600. The full log-prob rows come from the selection kernel (`store_all_logp`);
    the top-W expansion, parent bookkeeping and back-tracking are index arithmetic on the device (library top-k)."""

    name = "beam_search"

    def __init__(self, beam_width=None, select_best=True, **kwargs) -> None:
        kwargs["store_all_logp"] = True
        super().__init__(**kwargs)
        self.beam_width, self.select_best = beam_width, select_best
        self.cum, self.parents = None, []

    def pre_decoder_hook(self, td: TensorDict, env, action: torch.Tensor | None = None):
        """decoding.py:490-515: one beam per start node, forced first step with log-prob 0 and parent 0."""
        if self.beam_width is None:
            self.beam_width = env.get_num_starts(td)
        assert self.beam_width > 1, "beam width must be larger than 1"
        if self.select_start_nodes_fn is not None:
            action = self.select_start_nodes_fn(td, env, self.beam_width)
        else:
            action = env.select_start_nodes(td, num_starts=self.beam_width)
        td = batchify(td, self.beam_width)
        td.set("action", action)
        td = env.step(td)["next"]
        self.logprobs.append(torch.zeros_like(td["action_mask"], dtype=torch.float32))
        self.actions.append(action)
        self.parents.append(torch.zeros_like(action, dtype=torch.int32))
        self.cum = torch.zeros(action.shape[0], 1, dtype=torch.float32, device=action.device)
        return td, env, self.beam_width

    def step(self, logits: torch.Tensor, mask: torch.Tensor, td: TensorDict | None = None,
             action: torch.Tensor | None = None, **kwargs) -> TensorDict:
        assert td is not None, "td must be provided"
        self.actions, self.logprobs, kept_a, kept_lp = [], [], self.actions, self.logprobs
        td = super().step(logits, mask, td)  # one kernel: process_logits -> all log-probs (its arg-max is unused)
        all_lp = self.logprobs[0]
        self.actions, self.logprobs = kept_a, kept_lp
        node, parent, src, self.cum = beam_expand(all_lp, self.cum, self.beam_width)
        td = td[src]  # every surviving beam continues from its parent's state
        td.set("action", node)
        self.actions.append(node)
        self.logprobs.append(all_lp[src])
        self.parents.append(parent)
        return td

    def post_decoder_hook(self, td: TensorDict, env):
        actions, logprobs = beam_backtrack(torch.stack(self.actions, 1), torch.stack(self.logprobs, 1),
                                           torch.stack(self.parents, 1), self.beam_width)
        if self.select_best:  # decoding.py:558-566
            rewards = env.get_reward(td, actions)
            b = rewards.shape[0] // self.beam_width
            keep = torch.arange(b, device=rewards.device) + unbatchify(rewards, self.beam_width).argmax(dim=1) * b
            return logprobs[keep], actions[keep], td[keep], env
        return logprobs, actions, td, env


def get_decoding_strategy(decoding_strategy, **config) -> DecodingStrategy:
    """rl4co/utils/decoding.py:17-35 (an unknown name is an error here; the reference falls back to Sampling)."""
    registry = {"greedy": Greedy, "sampling": Sampling, "multistart_greedy": Greedy,
                "multistart_sampling": Sampling, "evaluate": Evaluate, "beam_search": BeamSearch}
    if decoding_strategy not in registry:
        raise NotImplementedError(f"decode type {decoding_strategy!r} is outside the fused path: {list(registry)}")
    if "multistart" in decoding_strategy:
        config["multistart"] = True
    return registry[decoding_strategy](**config)
