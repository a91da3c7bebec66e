"""Time + check co_encoder_mha variants against torch SDPA (fp64 reference).
Usage: B=65536 NS=100 VARIANTS=simt,tc,tc2 python tools/bench_mha.py   """
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from rl4co_b200 import native

B = int(os.environ.get('B', 65536))
VARIANTS = os.environ.get('VARIANTS', 'simt,tc,tc2,tc3').split(',')

def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def ref(qkv, b, n):
    x = qkv[: b * n].double().view(b, n, 3, 8, 16).permute(2, 0, 3, 1, 4)
    return F.scaled_dot_product_attention(x[0], x[1], x[2]).permute(0, 2, 1, 3).reshape(b * n, 128)

for N in [int(v) for v in os.environ.get('NS', '100,128,50,20,97').split(',')]:
    torch.manual_seed(N)
    qkv = torch.randn(B * N, 384, device="cuda") * float(os.environ.get('SCALE', 1.5))
    r = ref(qkv, 64, N)
    for v in VARIANTS:
        os.environ['CO_MHA_VARIANT'] = v
        out = native.encoder_mha(qkv, B, N)
        torch.cuda.synchronize()
        err = (out[: 64 * N].double() - r).abs().max().item()
        tail = (out[-64 * N:].double() - ref(qkv[-64 * N:], 64, N)).abs().max().item()
        print(f"N={N} variant={v}: {t(lambda: native.encoder_mha(qkv, B, N)):.2f} ms  max|err| head {err:.2e} tail {tail:.2e}", flush=True)
