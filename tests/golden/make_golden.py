"""Generate the golden vectors under tests/golden/ by executing rl4co's OWN, unmodified
hot-path files from /root/reference (through oracle/ref_standin.py -- container stubs for
tensordict/torchrl/lightning only, no arithmetic).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every array written here is produced by reference code + PyTorch CPU fp32.  The fixtures
pin (a) the env MDP (masks / visited / capacity / done / reward) under the reference's
``random_policy`` rollout, (b) the AM decoder path (raw logits per step, log-probs,
actions, reward, log-likelihood) for greedy, sampling-with-recorded-Exp(1)-noise,
multistart-greedy and teacher-forced evaluation, (c) the AM encoder, and (d) the
batchify / dihedral-8 layout conventions.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_standin  # noqa: E402

ref = ref_standin.load()
TensorDict = ref.TensorDict


def npy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def make_env(name, n, check=True):
    if name == "sdvrp":
        import importlib

        Env = importlib.import_module("rl4co.envs.routing.sdvrp.env").SDVRPEnv
    elif name == "op":
        import importlib

        Env = importlib.import_module("rl4co.envs.routing.op.env").OPEnv
    elif name == "pctsp":
        import importlib

        Env = importlib.import_module("rl4co.envs.routing.pctsp.env").PCTSPEnv
    else:
        Env = ref.TSPEnv if name == "tsp" else ref.CVRPEnv
    if name == "op":
        # the generator's default prize_distribution builds Uniform(1.0, 1.0), which current torch rejects at
        # construction; "dist" skips that sampler (it is unused: the prizes come from prize_type, op/generator.py:115-125)
        return Env(generator_params=dict(num_loc=n, prize_distribution="dist"), check_solution=check)
    return Env(generator_params=dict(num_loc=n), check_solution=check)


def env_fixture(name, n, batch, seed):
    """Reference env driven by the reference's random_policy (utils/decoding.py:78-106)."""
    torch.manual_seed(seed)
    env = make_env(name, n)
    td0 = env.generator(batch_size=[batch])
    out = {f"inst::{k}": npy(td0[k]) for k in td0.keys()}
    td = env.reset(td0.clone())
    masks, dones, visited, used, cur = [npy(td["action_mask"])], [], [], [], []
    actions = []
    while not td["done"].all():
        td = ref.decoding.random_policy(td)
        actions.append(td["action"].clone())
        td = env.step(td)["next"]
        masks.append(npy(td["action_mask"]))
        dones.append(npy(td["done"]))
        cur.append(npy(td["current_node"]).reshape(batch))
        if name == "cvrp":
            visited.append(npy(td["visited"]))
            used.append(npy(td["used_capacity"]))
        if name == "sdvrp":
            visited.append(npy(td["demand_with_depot"]))  # the dynamic state of the split-delivery env
            used.append(npy(td["used_capacity"]))
        if name == "op":
            visited.append(npy(td["visited"]))
            used.append(np.stack([npy(td["tour_length"]), npy(td["current_total_prize"])]))
        if name == "pctsp":
            visited.append(npy(td["visited"]))
            used.append(np.stack([npy(td["cur_total_prize"]), npy(td["cur_total_penalty"])]))
    actions = torch.stack(actions, 1)
    reward = env.get_reward(td, actions)  # runs check_solution_validity too
    out.update(actions=npy(actions), action_mask=np.stack(masks), done=np.stack(dones),
               current_node=np.stack(cur), reward=npy(reward))
    if name == "cvrp":
        out.update(visited=np.stack(visited), used_capacity=np.stack(used))
    elif name == "sdvrp":
        out.update(demand_with_depot=np.stack(visited), used_capacity=np.stack(used))
    elif name == "op":
        st = np.stack(used)  # [T, 2, B]
        out.update(visited=np.stack(visited), tour_length=st[:, 0], current_total_prize=st[:, 1])
    elif name == "pctsp":
        st = np.stack(used)
        out.update(visited=np.stack(visited), cur_total_prize=st[:, 0], cur_total_penalty=st[:, 1])
    else:
        out.update(first_node=npy(td["first_node"]), i=npy(td["i"]))
    return out


def sdvrp_am_fixture(n, batch, seed, name="sdvrp"):
    """A sibling env through the reference AttentionModelPolicy -- SDVRP (VRPContext + SDVRPDynamicEmbedding,
    dynamic.py:60-78) or OP (OPInitEmbedding init.py:254-280, OPContext context.py:201-213): greedy, sampling with
    recorded noise, teacher-forced evaluation -- single-start decoding."""
    torch.manual_seed(seed)
    env = make_env(name, n)
    pol = ref.AttentionModelPolicy(env_name=name, num_encoder_layers=1).eval()
    if name == "sdvrp":
        with torch.no_grad():  # the default init of a Linear(1, 3E) is small next to the static keys: make it count
            pol.decoder.dynamic_embedding.projection.weight.mul_(3.0)
    keep = ("decoder.",) if name == "sdvrp" else ("decoder.", "encoder.")
    out = {"w::" + k: npy(v) for k, v in pol.state_dict().items() if k.startswith(keep)}
    td0 = env.generator(batch_size=[batch])
    for k in td0.keys():
        out[f"inst::{k}"] = npy(td0[k])
    rec = Recorder(pol.decoder)
    with torch.inference_mode():
        td = env.reset(td0.clone())
        h, _ = pol.encoder(td)
        out["h"] = npy(h)
        o = pol(td.clone(), env, phase="test", decode_type="greedy", return_sum_log_likelihood=False)
        lg, mk = rec.pop()
        out.update(greedy_actions=npy(o["actions"]), greedy_logprobs=npy(o["log_likelihood"]),
                   greedy_reward=npy(o["reward"]), greedy_logits=lg, greedy_masks=mk)
        torch.manual_seed(seed + 1)
        o = pol(td.clone(), env, phase="train", decode_type="sampling", return_sum_log_likelihood=False)
        lg, mk = rec.pop()
        T = o["actions"].shape[1]
        torch.manual_seed(seed + 1)
        q = torch.stack([torch.empty(batch, lg.shape[-1]).exponential_(1) for _ in range(T)])
        out.update(sampling_actions=npy(o["actions"]), sampling_logprobs=npy(o["log_likelihood"]),
                   sampling_reward=npy(o["reward"]), sampling_noise=npy(q))
        torch.manual_seed(seed + 2)
        tdr = env.reset(td0.clone())
        acts = []
        while not tdr["done"].all():
            tdr = ref.decoding.random_policy(tdr)
            acts.append(tdr["action"].clone())
            tdr = env.step(tdr)["next"]
        acts = torch.stack(acts, 1)
        o = pol(td.clone(), env, phase="train", actions=acts, return_sum_log_likelihood=False)
        lg, mk = rec.pop()
        out.update(eval_actions=npy(acts), eval_logprobs=npy(o["log_likelihood"]), eval_reward=npy(o["reward"]),
                   eval_logits=lg)
    return out


class Recorder:
    """Forward hook on the reference decoder recording raw (logits, mask) per step."""

    def __init__(self, decoder):
        self.logits, self.masks = [], []
        self.h = decoder.register_forward_hook(self._hook)

    def _hook(self, mod, args, output):
        self.logits.append(output[0].detach().clone())
        self.masks.append(output[1].detach().clone())

    def pop(self):
        lg, mk = torch.stack(self.logits), torch.stack(self.masks)
        self.logits, self.masks = [], []
        return npy(lg), npy(mk)


def am_fixture(name, n, batch, seed, ms_batch=3, lean=False):
    """`lean` drops the recorded sampling / multistart logits (4 MB at N = 100); the greedy and
    teacher-forced logits, every action / log-prob / reward and the consumed noise stay."""
    torch.manual_seed(seed)
    env = make_env(name, n)
    pol = ref.AttentionModelPolicy(env_name=name, num_encoder_layers=1).eval()
    # make the decoder more "opinionated" than default init so that argmax margins are healthy
    out = {}
    for k, v in pol.state_dict().items():
        if k.startswith("decoder."):
            out["w::" + k] = npy(v)
    td0 = env.generator(batch_size=[batch])
    for k in td0.keys():
        out[f"inst::{k}"] = npy(td0[k])
    rec = Recorder(pol.decoder)
    with torch.inference_mode():
        td = env.reset(td0.clone())
        h, _ = pol.encoder(td)
        out["h"] = npy(h)
        # -- greedy
        o = pol(td.clone(), env, phase="test", decode_type="greedy", return_sum_log_likelihood=False)
        lg, mk = rec.pop()
        out.update(greedy_actions=npy(o["actions"]), greedy_logprobs=npy(o["log_likelihood"]),
                   greedy_reward=npy(o["reward"]), greedy_logits=lg, greedy_masks=mk)
        # -- sampling; the Exp(1) draws consumed by torch.multinomial are regenerated
        torch.manual_seed(seed + 1)
        o = pol(td.clone(), env, phase="train", decode_type="sampling", return_sum_log_likelihood=False)
        lg, mk = rec.pop()
        T = o["actions"].shape[1]
        torch.manual_seed(seed + 1)
        q = torch.stack([torch.empty(batch, lg.shape[-1]).exponential_(1) for _ in range(T)])
        out.update(sampling_actions=npy(o["actions"]), sampling_logprobs=npy(o["log_likelihood"]),
                   sampling_reward=npy(o["reward"]), sampling_noise=npy(q))
        if not lean:
            out.update(sampling_logits=lg)
        # -- teacher-forced evaluation of an independent (random-policy) action sequence
        torch.manual_seed(seed + 2)
        tdr = env.reset(td0.clone())
        acts = []
        while not tdr["done"].all():
            tdr = ref.decoding.random_policy(tdr)
            acts.append(tdr["action"].clone())
            tdr = env.step(tdr)["next"]
        acts = torch.stack(acts, 1)
        o = pol(td.clone(), env, phase="train", actions=acts, return_sum_log_likelihood=False)
        lg, mk = rec.pop()
        out.update(eval_actions=npy(acts), eval_logprobs=npy(o["log_likelihood"]), eval_reward=npy(o["reward"]),
                   eval_logits=lg)
        # -- multistart greedy on the first ms_batch instances (start-major layout)
        tdm = env.reset(td0[:ms_batch].clone())
        o = pol(tdm.clone(), env, phase="test", decode_type="multistart_greedy", return_sum_log_likelihood=False)
        lg, mk = rec.pop()
        out.update(ms_actions=npy(o["actions"]), ms_logprobs=npy(o["log_likelihood"]), ms_reward=npy(o["reward"]),
                   ms_batch=np.int64(ms_batch))
        if not lean:
            out.update(ms_logits=lg)
        # -- POMO-style: no graph context (pomo/model.py:59-63)
        polp = ref.AttentionModelPolicy(env_name=name, num_encoder_layers=1, use_graph_context=False).eval()
        polp.load_state_dict(pol.state_dict())
        o = polp(tdm.clone(), env, phase="test", decode_type="multistart_greedy", return_sum_log_likelihood=False)
        out.update(pomo_actions=npy(o["actions"]), pomo_logprobs=npy(o["log_likelihood"]), pomo_reward=npy(o["reward"]))
    return out


def decoding_fixture(name, n, batch, seed):
    """Beam search and top-k / top-p sampling (utils/decoding.py:109-188,464-600) on the SAME policy and instances
    as am_fixture(name, n, batch, seed): the seed sequence up to the generator call is replayed, weights and
    instances are not stored again (`h` is, as a guard)."""
    torch.manual_seed(seed)
    env = make_env(name, n)
    pol = ref.AttentionModelPolicy(env_name=name, num_encoder_layers=1).eval()
    td0 = env.generator(batch_size=[batch])
    out = {}
    with torch.inference_mode():
        td = env.reset(td0.clone())
        h, _ = pol.encoder(td)
        out["h"] = npy(h)
        for tag, kw in (("beam3_best", dict(beam_width=3, select_best=True)),
                        ("beam4_all", dict(beam_width=4, select_best=False)),
                        ("beamN_best", dict(select_best=True))):
            o = pol(td.clone(), env, phase="test", decode_type="beam_search", return_sum_log_likelihood=False, **kw)
            out.update({f"{tag}_actions": npy(o["actions"]), f"{tag}_logprobs": npy(o["log_likelihood"]),
                        f"{tag}_reward": npy(o["reward"])})
        for i, (tag, kw) in enumerate((("topk4", dict(top_k=4)), ("topp80", dict(top_p=0.8)),
                                       ("topk6_topp90", dict(top_k=6, top_p=0.9)))):
            torch.manual_seed(seed + 10 + i)
            o = pol(td.clone(), env, phase="train", decode_type="sampling", return_sum_log_likelihood=False, **kw)
            T = o["actions"].shape[1]
            torch.manual_seed(seed + 10 + i)  # the Exp(1) draws torch.multinomial consumed, one [B, N] block per step
            q = torch.stack([torch.empty(batch, td["action_mask"].shape[-1]).exponential_(1) for _ in range(T)])
            out.update({f"{tag}_actions": npy(o["actions"]), f"{tag}_logprobs": npy(o["log_likelihood"]),
                        f"{tag}_reward": npy(o["reward"]), f"{tag}_noise": npy(q)})
    return out


def encoder_fixture(name, n, batch, seed, normalization):
    torch.manual_seed(seed)
    env = make_env(name, n)
    pol = ref.AttentionModelPolicy(env_name=name, num_encoder_layers=1, normalization=normalization).eval()
    if normalization == "batch":  # non-trivial running stats
        for m in pol.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    out = {"w::" + k: npy(v) for k, v in pol.state_dict().items()
           if k.startswith("encoder.") and "num_batches" not in k}
    td0 = env.generator(batch_size=[batch])
    for k in td0.keys():
        out[f"inst::{k}"] = npy(td0[k])
    with torch.inference_mode():
        td = env.reset(td0.clone())
        h, init_h = pol.encoder(td)
    out.update(h=npy(h), init_h=npy(init_h))
    return out


def name_seeded_weights(module, seed):
    """Deterministic weights that depend only on (parameter name, shape, seed) -- not on module
    construction order -- so the consumer regenerates them instead of the fixture storing 5 MB:
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for matrices (torch.nn.Linear's range), U(-.1,.1) for
    vectors, norm weights 1 + U(-.1,.1).  Mirrored by tests/conftest.py::name_seeded_weights."""
    import zlib

    sd = module.state_dict()
    new = {}
    for k in sorted(sd.keys()):
        v = sd[k]
        if not v.dtype.is_floating_point:
            new[k] = v.clone()
            continue
        gen = torch.Generator().manual_seed(seed * 1_000_003 + zlib.crc32(k.encode()))
        u = torch.rand(v.shape, generator=gen) * 2 - 1
        if v.dim() >= 2:
            new[k] = u / (v.shape[-1] ** 0.5)
        elif "W_placeholder" in k:
            new[k] = u
        elif "norm" in k and k.endswith("weight"):
            new[k] = 1 + 0.1 * u
        elif k.endswith("running_var"):
            new[k] = 1 + 0.25 * u
        else:
            new[k] = 0.1 * u
    module.load_state_dict(new)
    return new


def pomo_fixture(n, batch, seed, num_augment=8):
    """BASELINE config C4 at its own scale: TSP-n POMO = 6-layer instance-norm encoder, no graph
    context (zoo/pomo/model.py:59-63), dihedral-8 StateAugmentation (aug-major, data/transforms.py:
    16-38,118-151), multistart greedy with n starts (start-major), then POMO's reductions
    (pomo/model.py:103-136): max over starts, max over augmentations."""
    torch.manual_seed(seed)
    env = make_env("tsp", n, check=True)
    pol = ref.AttentionModelPolicy(env_name="tsp", num_encoder_layers=6, normalization="instance",
                                   use_graph_context=False).eval()
    name_seeded_weights(pol, seed)
    td0 = env.generator(batch_size=[batch])
    out = {"inst::locs": npy(td0["locs"]), "weight_seed": np.int64(seed), "num_augment": np.int64(num_augment)}
    with torch.inference_mode():
        td = env.reset(td0.clone())
        tda = ref.transforms.StateAugmentation(num_augment=num_augment, augment_fn="dihedral8")(td)
        out["aug_locs"] = npy(tda["locs"])
        h, _ = pol.encoder(tda)
        out["h_first_rows"] = npy(h[:2])  # encoder guard: aug 0 of the first two instances
        o = pol(tda.clone(), env, phase="test", decode_type="multistart_greedy", num_starts=n,
                return_sum_log_likelihood=False)
        acts, lp, rew = o["actions"], o["log_likelihood"], o["reward"]
        assert acts.max() < 256
        r = ref.ops.unbatchify(rew, (num_augment, n))            # [B, aug, start]
        max_r, _ = r.max(dim=-1)
        max_aug_r, _ = max_r.max(dim=1)
        out.update(actions=npy(acts).astype(np.uint8), logprobs_sum=npy(lp.sum(1)), reward=npy(rew),
                   logprobs_rows=npy(lp[: 2 * batch]),  # full per-step log-probs of the first rows
                   max_reward=npy(max_r), max_aug_reward=npy(max_aug_r), reward_b_aug_start=npy(r))
    return out


def layout_fixture():
    torch.manual_seed(11)
    x = torch.rand(3, 5, 2)
    aug = ref.transforms.dihedral_8_augmentation(x)
    td = TensorDict({"locs": x.clone()}, batch_size=[3])
    sa = ref.transforms.StateAugmentation(num_augment=8, augment_fn="dihedral8")(td)
    b = ref.ops.batchify(x, 4)
    r = torch.arange(3 * 8 * 4, dtype=torch.float32)
    ub = ref.ops.unbatchify(r, (8, 4))
    env = make_env("tsp", 5)
    envc = make_env("cvrp", 5)
    tdx = env.reset(batch_size=[3])
    tdc = envc.reset(batch_size=[3])
    return dict(x=npy(x), dihedral8=npy(aug), state_aug=npy(sa["locs"]), batchify4=npy(b),
                unbatchify_8_4=npy(ub), tsp_starts=npy(env.select_start_nodes(tdx, 5)),
                cvrp_starts=npy(envc.select_start_nodes(tdc, 5)),
                tsp_num_starts=np.int64(env.get_num_starts(tdx)), cvrp_num_starts=np.int64(envc.get_num_starts(tdc)))


def main():
    jobs = {
        "env_tsp20": lambda: env_fixture("tsp", 20, 16, 100),
        "env_tsp50": lambda: env_fixture("tsp", 50, 8, 101),
        "env_cvrp20": lambda: env_fixture("cvrp", 20, 16, 102),
        "env_cvrp50": lambda: env_fixture("cvrp", 50, 8, 103),
        "am_tsp20": lambda: am_fixture("tsp", 20, 8, 200),
        "am_cvrp20": lambda: am_fixture("cvrp", 20, 8, 201),
        "am_tsp50": lambda: am_fixture("tsp", 50, 4, 202, ms_batch=2),
        "am_cvrp50": lambda: am_fixture("cvrp", 50, 4, 203, ms_batch=2),
        "am_tsp100": lambda: am_fixture("tsp", 100, 4, 204, ms_batch=1, lean=True),
        "am_cvrp100": lambda: am_fixture("cvrp", 100, 4, 205, ms_batch=1, lean=True),
        "pomo_tsp100": lambda: pomo_fixture(100, 2, 400),
        "env_sdvrp20": lambda: env_fixture("sdvrp", 20, 16, 104),
        "env_sdvrp50": lambda: env_fixture("sdvrp", 50, 8, 105),
        "am_sdvrp20": lambda: sdvrp_am_fixture(20, 8, 206),
        "am_sdvrp50": lambda: sdvrp_am_fixture(50, 4, 207),
        "env_op20": lambda: env_fixture("op", 20, 16, 108),
        "env_op50": lambda: env_fixture("op", 50, 8, 109),
        "am_op20": lambda: sdvrp_am_fixture(20, 8, 210, name="op"),
        "am_op50": lambda: sdvrp_am_fixture(50, 4, 211, name="op"),
        "env_pctsp20": lambda: env_fixture("pctsp", 20, 16, 110),
        "env_pctsp50": lambda: env_fixture("pctsp", 50, 8, 111),
        "am_pctsp20": lambda: sdvrp_am_fixture(20, 8, 212, name="pctsp"),
        "am_pctsp50": lambda: sdvrp_am_fixture(50, 4, 213, name="pctsp"),
        "enc_tsp20_batch": lambda: encoder_fixture("tsp", 20, 4, 300, "batch"),
        "enc_cvrp20_instance": lambda: encoder_fixture("cvrp", 20, 4, 301, "instance"),
        "layout": layout_fixture,
        "dec_tsp20": lambda: decoding_fixture("tsp", 20, 8, 200),
        "dec_cvrp20": lambda: decoding_fixture("cvrp", 20, 8, 201),
    }
    only = set(sys.argv[1:])  # `python make_golden.py dec_tsp20 dec_cvrp20` regenerates just those files
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        data = fn()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {len(data)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
