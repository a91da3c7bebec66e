"""Measure the BASELINE.json configs C2-C5 (single-GPU share) on cuda:0 -> JSON lines.
selections = exact count of (decode -> select -> env.step) iterations (sum of per-trajectory steps)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl4co_b200.envs import get_env
from rl4co_b200.policy import FusedAttentionModelPolicy
from rl4co_b200.reinforce import get_reinforce_baseline, pomo_step, reinforce_step

dev = torch.device("cuda:0")


def timed(fn, warm=2, it=4):
    for _ in range(warm):
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it, r


def policy_for(env_name, **kw):
    torch.manual_seed(0)
    return FusedAttentionModelPolicy(env_name=env_name, **kw).to(dev).eval()


def run(name, env_name, n, B, decode_type, **kw):
    env = get_env(env_name, generator_params=dict(num_loc=n), check_solution=False)
    pol = policy_for(env_name)
    torch.manual_seed(1234)
    td = env.reset(env.generator(B).to(dev))
    with torch.inference_mode():
        ms, out = timed(lambda: pol(td, env, phase="test", decode_type=decode_type, **kw))
        enc = pol.encoder(td)
        ms_dec, _ = timed(lambda: pol(td, env, phase="test", decode_type=decode_type, encoder_output=enc, **kw))
    nsel = out["actions"].numel()
    print(json.dumps({"config": name, "batch": B, "T": out["actions"].shape[1], "selections": nsel,
                      "policy_forward_ms": ms, "policy_forward_sel_per_s": nsel / ms * 1e3,
                      "decode_only_ms": ms_dec, "decode_only_sel_per_s": nsel / ms_dec * 1e3,
                      "reward_mean": out["reward"].mean().item()}))


run("C1 TSP-20 greedy B=128", "tsp", 20, 128, "greedy")
run("C2 TSP-50 greedy B=4096", "tsp", 50, 4096, "greedy")
run("C3 CVRP-50 sampling B=4096", "cvrp", 50, 4096, "sampling", seed=1)
run("TSP-100 greedy B=4096", "tsp", 100, 4096, "greedy")
run("CVRP-100 sampling B=8192", "cvrp", 100, 8192, "sampling", seed=1)

# C4: TSP-100 POMO, 8 aug x 100 starts; per-GPU share of 1024 instances = 128
env = get_env("tsp", generator_params=dict(num_loc=100), check_solution=False)
pol = policy_for("tsp", num_encoder_layers=6, normalization="instance", use_graph_context=False)
torch.manual_seed(1234)
td = env.reset(env.generator(128).to(dev))
ms, res = timed(lambda: pomo_step(pol, env, td, num_augment=8, phase="test"), warm=1, it=3)
nsel = res["actions"].numel()
print(json.dumps({"config": "C4 TSP-100 POMO 8aug x 100 starts, 128 instances (1/8 of 1024)", "trajectories": res["actions"].shape[0],
                  "selections": nsel, "ms": ms, "sel_per_s": nsel / ms * 1e3, "max_aug_reward_mean": res["max_aug_reward"].mean().item()}))

# C5: CVRP-100 REINFORCE step, per-GPU share of 65536 = 8192 instances
env = get_env("cvrp", generator_params=dict(num_loc=100), check_solution=False)
pol = policy_for("cvrp")
opt = torch.optim.Adam(pol.parameters(), lr=1e-4)
bl = get_reinforce_baseline("mean")
torch.manual_seed(1234)
td = env.reset(env.generator(8192).to(dev))
ms, res = timed(lambda: reinforce_step(pol, env, td, bl, opt, seed=5), warm=1, it=3)
nsel = res["actions"].numel()
with torch.inference_mode():
    pol.eval()
    ms_roll, out = timed(lambda: pol(td, env, decode_type="sampling", seed=5), warm=1, it=3)
print(json.dumps({"config": "C5 CVRP-100 REINFORCE step, 8192 instances (1/8 of 65536)", "T": res["actions"].shape[1],
                  "selections": nsel, "train_step_ms": ms, "train_step_sel_per_s": nsel / ms * 1e3,
                  "rollout_policy_forward_ms": ms_roll, "rollout_sel_per_s": out["actions"].numel() / ms_roll * 1e3,
                  "loss": res["loss"].item()}))
