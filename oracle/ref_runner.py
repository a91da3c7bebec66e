"""TEST / BENCH INFRASTRUCTURE ONLY -- times the reference's OWN policy on the host CPU cores.

Used by ``bench.py --impl reference`` and by ``bench.py``'s ``cpu_baseline`` leg.  Runs the
unmodified ``rl4co`` ``AttentionModelPolicy`` / ``TSPEnv`` / ``CVRPEnv`` files (from
``/root/reference`` in the build container, from the staged ``oracle/_ref`` copy on the GPU box,
see ``oracle/make_ref.py``) through the container stubs of ``oracle/ref_standin.py``:
``policy(env.reset(td), env, phase="test", decode_type=...)`` -- encoder + decode loop + reward,
fp32, ``torch.inference_mode()``, ``check_solution=False`` (SURVEY.md 8d "CPU baseline timing").

Two layouts are timed so that "all the host threads it can use" is answered honestly:
  * one process, the best intra-op thread count of a small probe (torch CPU ops on [B,N,128]
    tensors stop scaling -- and regress -- well below a 100+-core host's width);
  * P processes x T threads covering every core (P*T = cores), each with its own batch.
The reported value is the better of the two (whole-host selections per second).

When neither reference tree exists the oracle port (``am_rollout_oracle``) is timed instead and
the result says ``kind: "port"``.
"""

from __future__ import annotations

import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _policy_kwargs(workload_kwargs):
    kw = dict(embed_dim=128, num_heads=8, num_encoder_layers=3, normalization="batch", tanh_clipping=10.0)
    kw.update(workload_kwargs or {})
    return kw


def _build(env_name, num_loc, policy_kwargs, augment=0, train=False):
    """-> (kind, make_batch, run(batch) -> selections) using the real reference when available.
    `augment` = 8: dihedral-8 StateAugmentation before the policy (POMO val/test, pomo/model.py:97-101);
    `train`: policy-forward in train mode with autograd + REINFORCE loss (mean baseline) + backward + Adam step
    (reinforce.py:59-111; the Lightning module around it is not needed for the arithmetic)."""
    import torch

    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_standin

    if ref_standin.reference_available():
        ref = ref_standin.load()
        Env = ref.TSPEnv if env_name == "tsp" else ref.CVRPEnv
        env = Env(generator_params=dict(num_loc=num_loc), check_solution=False)
        torch.manual_seed(0)
        policy = ref.AttentionModelPolicy(env_name=env_name, **policy_kwargs)
        policy = policy.train() if train else policy.eval()
        opt = torch.optim.Adam(policy.parameters(), lr=1e-4) if train else None
        aug = ref.transforms.StateAugmentation(num_augment=augment, augment_fn="dihedral8") if augment else None

        def make_batch(batch, seed):
            torch.manual_seed(seed)
            return env.generator(batch_size=[batch])

        def run(td0, decode_type, **kw):
            if train:
                out = policy(env.reset(td0.clone()), env, phase="train", decode_type=decode_type, **kw)
                loss = -((out["reward"] - out["reward"].mean()) * out["log_likelihood"]).mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                return out["actions"].numel()
            with torch.inference_mode():
                td = env.reset(td0.clone())
                if aug is not None:
                    td = aug(td)
                out = policy(td, env, phase="test", decode_type=decode_type, **kw)
            return out["actions"].numel()

        return "reference", make_batch, run

    from oracle import am_rollout_oracle as O
    from rl4co_b200.policy import FusedAttentionModelPolicy

    torch.manual_seed(0)
    pol = FusedAttentionModelPolicy(env_name=env_name, **policy_kwargs).eval()
    W = {k: v.detach() for k, v in pol.state_dict().items()}
    nl = policy_kwargs.get("num_encoder_layers", 3)

    def make_batch(batch, seed):
        torch.manual_seed(seed)
        return O.generate_instances(env_name, batch, num_loc)

    def run(inst, decode_type, **kw):
        if augment or train:
            raise RuntimeError("the oracle-port fallback only times plain policy-forward workloads")
        with torch.inference_mode():
            out = O.policy_forward(W, env_name, inst, num_layers=nl, decode_type=decode_type, faithful_copies=True,
                                   normalization=policy_kwargs.get("normalization", "batch"))
        return out["actions"].numel()

    return "port", make_batch, run


def worker(env_name, num_loc, batch, decode_type, steps, warmup, threads, seed, policy_kwargs, decode_kwargs,
           augment=0, train=False):
    """One process: `warmup` untimed + `steps` timed policy-forward calls. Returns a dict."""
    import torch

    torch.set_num_threads(threads)
    kind, make_batch, run = _build(env_name, num_loc, _policy_kwargs(policy_kwargs), augment, train)
    td0 = make_batch(batch, seed)
    for _ in range(warmup):
        run(td0, decode_type, **(decode_kwargs or {}))
    t_start = time.time()
    per = []
    nsel = 0
    for _ in range(steps):
        t0 = time.perf_counter()
        nsel += run(td0, decode_type, **(decode_kwargs or {}))
        per.append(time.perf_counter() - t0)
    return {"kind": kind, "selections": nsel, "t_start": t_start, "t_end": time.time(), "step_s": per,
            "threads": threads}


def _spawn(env_name, num_loc, batch, decode_type, steps, warmup, threads, seed, policy_kwargs, decode_kwargs,
           augment=0, train=False):
    spec = json.dumps(dict(env_name=env_name, num_loc=num_loc, batch=batch, decode_type=decode_type, steps=steps,
                           warmup=warmup, threads=threads, seed=seed, policy_kwargs=policy_kwargs,
                           decode_kwargs=decode_kwargs, augment=augment, train=train))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads), CUDA_VISIBLE_DEVICES="")
    return subprocess.Popen([sys.executable, os.path.abspath(__file__), spec], stdout=subprocess.PIPE,
                            stderr=subprocess.DEVNULL, text=True, env=env, cwd=ROOT)


def _collect(procs):
    outs = []
    for p in procs:
        txt, _ = p.communicate()
        line = [ln for ln in txt.splitlines() if ln.startswith("{")]
        if p.returncode != 0 or not line:
            raise RuntimeError(f"reference worker failed (rc={p.returncode})")
        outs.append(json.loads(line[-1]))
    return outs


def time_reference(env_name, num_loc, batch, decode_type, steps=2, warmup=1, policy_kwargs=None,
                   decode_kwargs=None, multi_process=True, augment=0, train=False):
    """Whole-host throughput of the reference policy-forward. Returns a dict for bench.py."""
    cores = os.cpu_count() or 1
    # (1) one process: probe the intra-op thread count on a small batch
    cand = sorted({min(c, cores) for c in (8, 16, 32, 64)})
    best_t, best_rate = cand[0], 0.0
    for nt in cand:
        r = _collect([_spawn(env_name, num_loc, min(128, batch), decode_type, 1, 1, nt, 99, policy_kwargs,
                             decode_kwargs, augment, train)])[0]
        rate = r["selections"] / sum(r["step_s"])
        if rate > best_rate:
            best_t, best_rate = nt, rate
    single = _collect([_spawn(env_name, num_loc, batch, decode_type, steps, warmup, best_t, 1234, policy_kwargs,
                              decode_kwargs, augment, train)])[0]
    single_rate = single["selections"] / sum(single["step_s"])
    res = {"kind": single["kind"], "host_cores": cores, "single": {"threads": best_t, "value": single_rate,
                                                                 "ms_per_step": 1e3 * sum(single["step_s"]) / steps,
                                                                 "batch": batch}}
    # (2) P processes x T threads over every core
    if multi_process and cores >= 16:
        T = 16 if cores >= 32 else 8
        P = max(1, cores // T)
        # bounded: half the per-process sample and at most two timed passes -- on a loaded 128-core host one pass of
        # 8 x 1 024 TSP-100 instances took 51 s, and the leg must stay within "10-30 s of CPU work per measurement"
        m_batch, m_steps = max(min(batch, 128), batch // 2), min(steps, 2)
        procs = [_spawn(env_name, num_loc, m_batch, decode_type, m_steps, warmup, T, 1234 + i, policy_kwargs,
                        decode_kwargs, augment, train) for i in range(P)]
        outs = _collect(procs)
        # every worker runs the same amount of work concurrently: aggregate = total / slowest worker's timed span
        span = max(sum(o["step_s"]) for o in outs)
        res["multi"] = {"processes": P, "threads_each": T, "value": sum(o["selections"] for o in outs) / span,
                        "ms_per_step": 1e3 * span / m_steps, "batch": m_batch * P}
    best = res["single"]
    res["layout"] = f"1 process x {best_t} threads"
    res["cores_used"] = best_t
    if "multi" in res and res["multi"]["value"] > best["value"]:
        best = res["multi"]
        res["layout"] = f"{res['multi']['processes']} processes x {res['multi']['threads_each']} threads"
        res["cores_used"] = res["multi"]["processes"] * res["multi"]["threads_each"]
    res["value"], res["ms_per_step"], res["batch_per_step"] = best["value"], best["ms_per_step"], best["batch"]
    return res


if __name__ == "__main__":
    spec = json.loads(sys.argv[1])
    print(json.dumps(worker(**spec)))
