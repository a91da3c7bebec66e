"""Sibling envs: policy-forward time of SDVRP / OP through the persistent kernel and through the stepping
kernels on the same instances (sampling, so that the untrained policy produces long tours).
usage: python tools/bench_sibling_envs.py [B] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rl4co_b200.envs import get_env
from rl4co_b200.policy import FusedAttentionModelPolicy

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")


def t(fn, n=3):
    out = fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


for env_name in ("cvrp", "sdvrp", "op", "pctsp"):
    torch.manual_seed(0)
    env = get_env(env_name, generator_params=dict(num_loc=N), check_solution=False)
    pol = FusedAttentionModelPolicy(env_name=env_name).to(dev).eval()
    td_host = env.generator(B)
    with torch.inference_mode():
        fused_ms, out = t(lambda: pol(env.reset(td_host.to(dev)), env, phase="test", decode_type="sampling", seed=1))
        step_ms, out2 = t(lambda: pol(env.reset(td_host.to(dev)), env, phase="test", decode_type="sampling",
                                      fused_rollout=False), n=1)
    sel = out["actions"].numel()
    print(f"{env_name.upper()}-{N} sampling B={B}: persistent kernel {fused_ms:.2f} ms ({sel / fused_ms * 1e3:.3e} selections/s, "
          f"T={out['actions'].shape[1]}), stepping kernels {step_ms:.1f} ms (T={out2['actions'].shape[1]}), x{step_ms / fused_ms:.1f}")
