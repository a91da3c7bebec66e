"""REINFORCE training step of the budget envs (orienteering, prize-collecting TSP) next to CVRP: fused sampling rollout +
one-call teacher-forced pass (reinforce.replay_budget_states, co_attn_fwd / co_attn_bwd) + backward + Adam.
usage: python tools/bench_budget_train.py [B] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rl4co_b200.envs import get_env
from rl4co_b200.policy import FusedAttentionModelPolicy
from rl4co_b200.reinforce import get_reinforce_baseline, reinforce_step

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
for env_name in ("cvrp", "sdvrp", "op", "pctsp"):
    torch.manual_seed(0)
    env = get_env(env_name, generator_params=dict(num_loc=N), check_solution=False)
    pol = FusedAttentionModelPolicy(env_name=env_name).to(dev)
    opt = torch.optim.Adam(pol.parameters(), lr=1e-4)
    bl = get_reinforce_baseline("exponential")
    td = env.reset(env.generator(B).to(dev))
    step = lambda s: reinforce_step(pol, env, td, bl, optimizer=opt, seed=s, matmul_precision="medium")
    for s in range(3):
        res = step(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(3, 8):
        res = step(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{env_name.upper()}-{N} REINFORCE step B={B}: {ms:.2f} ms ({B / ms * 1e3:.3e} instances/s, T={res['actions'].shape[1]}, "
          f"reward {res['reward'].mean().item():.3f}, loss {res['loss'].item():.4f})")
