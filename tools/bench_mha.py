import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl4co_b200 import native
B, N = int(os.environ.get('B', 65536)), 100
qkv = torch.randn(B * N, 384, device="cuda")
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"CO_MHA_VARIANT={os.environ.get('CO_MHA_VARIANT','1')}: {t(lambda: native.encoder_mha(qkv, B, N)):.2f} ms")
