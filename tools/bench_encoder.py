"""Time the fused encoder path under a few settings (chunk size, split-K for the K=512 GEMM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl4co_b200.envs import get_env
from rl4co_b200.policy import FusedAttentionModelPolicy
import rl4co_b200.encoder as encmod

dev = torch.device("cuda:0")
torch.manual_seed(0)
pol = FusedAttentionModelPolicy(env_name="tsp").to(dev).eval()
env = get_env("tsp", generator_params=dict(num_loc=100))
td = env.reset(env.generator(65536).to(dev))

def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

orig = encmod.AttentionModelEncoder._linear_splitk
for chunk in (8192, 16384, 65536):
    for splitk in (True, False):
        pol.encoder.inference_chunk = chunk
        if splitk:
            encmod.AttentionModelEncoder._linear_splitk = orig
        else:
            def nosplit(self, x, lin, residual, aff):
                from rl4co_b200 import native
                hi, lo = self._split(lin.weight)
                return native.gemm_tf32x3(x, hi, lo, bias=lin.bias, residual=residual,
                                          scale=aff[0] if aff is not None else None, shift=aff[1] if aff is not None else None)
            encmod.AttentionModelEncoder._linear_splitk = nosplit
        with torch.inference_mode():
            ms = t(lambda: pol.encoder(td))
        print(f"chunk {chunk:6d} splitk {splitk}: encoder {ms:7.1f} ms")
