"""Small driver for ncu: one no-grad encoder forward (3 layers, batch-norm) so that co::ffn::ffn_fused_kernel,
the attention kernel and the GEMM pipe kernel can each be captured once.  usage: python tools/profile_encoder.py [B] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rl4co_b200.envs import get_env
from rl4co_b200.policy import FusedAttentionModelPolicy

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
torch.manual_seed(0)
pol = FusedAttentionModelPolicy(env_name="tsp", num_encoder_layers=3).to(dev).eval()
env = get_env("tsp", generator_params=dict(num_loc=N), check_solution=False)
with torch.inference_mode():
    td = env.reset(env.generator(B).to(dev))
    for _ in range(2):
        h, _ = pol.encoder(td)
torch.cuda.synchronize()
print("encoder output", tuple(h.shape), float(h.abs().mean()))
