"""Per-kernel CUDA time of one POMO evaluation step (BASELINE config C4: dihedral-8 x 100 starts, 6-layer instance-norm
encoder, no graph context) -- torch.profiler kernel table.  usage: python tools/time_pomo_kernels.py [B] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from rl4co_b200.envs import get_env
from rl4co_b200.policy import FusedAttentionModelPolicy
from rl4co_b200.reinforce import pomo_step

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
torch.manual_seed(0)
pol = FusedAttentionModelPolicy(env_name="tsp", num_encoder_layers=6, normalization="instance", use_graph_context=False).to(dev).eval()
env = get_env("tsp", generator_params=dict(num_loc=N), check_solution=False)
td = env.reset(env.generator(B).to(dev))
for _ in range(2):
    out = pomo_step(pol, env, td, num_augment=8, phase="test")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = pomo_step(pol, env, td, num_augment=8, phase="test")
e1.record()
torch.cuda.synchronize()
print(f"pomo_step B={B} N={N}: {e0.elapsed_time(e1):.2f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = pomo_step(pol, env, td, num_augment=8, phase="test")
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90))
