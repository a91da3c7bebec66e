"""Per-kernel CUDA time of one no-grad encoder forward + cache GEMM + rollout (torch.profiler / CUPTI kernel activity; the
library's kernels are launched through ctypes but show up by name).  usage: python tools/time_encoder_kernels.py [B] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from rl4co_b200.envs import get_env
from rl4co_b200.policy import FusedAttentionModelPolicy

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
env_name = os.environ.get("ENV", "tsp")
dev = torch.device("cuda:0")
torch.manual_seed(0)
pol = FusedAttentionModelPolicy(env_name=env_name).to(dev).eval()
env = get_env(env_name, generator_params=dict(num_loc=N), check_solution=False)
with torch.inference_mode():
    td = env.reset(env.generator(B).to(dev))
    for _ in range(2):
        out = pol(td, env, decode_type="greedy")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = pol(td, env, decode_type="greedy")
    e1.record()
    torch.cuda.synchronize()
    print(f"policy forward {env_name} B={B} N={N}: {e0.elapsed_time(e1):.2f} ms")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        out = pol(td, env, decode_type="greedy")
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
