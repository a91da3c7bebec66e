"""CUDA-event times of co_attn_fwd / co_attn_bwd (dQ + dK/dV) at the training-chunk sizes, next to torch's SDPA
forward / backward on the same tensors.  usage: python tools/bench_attn_train.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from rl4co_b200 import attention_train as AT
from rl4co_b200 import native

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for M, N, masked in ((101, 101, False), (131, 101, True)):
    torch.manual_seed(0)
    q = torch.randn(B, M, 128, device=dev)
    cache = torch.randn(B, N, 512, device=dev)
    k, v = cache[..., :128], cache[..., 128:256]
    mask = (torch.rand(B, M, N, device=dev) < 0.6) if masked else None
    if masked:
        mask[..., 0] = True
    mw = AT.pack_mask(mask) if masked else None
    o = torch.empty(B, M, 128, device=dev); lse = torch.empty(B, 8, M, device=dev)
    dO = torch.randn_like(o); dq = torch.empty_like(q); dk = torch.empty(B, N, 128, device=dev); dv = torch.empty_like(dk)
    f = t(lambda: native.attn_fwd(q, k, v, mw, o, lse))
    b = t(lambda: native.attn_bwd(q, k, v, mw, o, lse, dO, dq, dk, dv))
    pm = t(lambda: AT.pack_mask(mask)) if masked else 0.0

    def heads(x):
        return x.reshape(B, x.shape[1], 8, 16).transpose(1, 2)

    qh, kh, vh = heads(q).requires_grad_(True), heads(k).detach().requires_grad_(True), heads(v).detach().requires_grad_(True)
    am = mask[:, None] if masked else None
    sf = t(lambda: F.scaled_dot_product_attention(qh, kh, vh, attn_mask=am))
    out = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=am)
    g = torch.randn_like(out)
    sb = t(lambda: torch.autograd.grad(out, (qh, kh, vh), g, retain_graph=True))
    print(f"B={B} M={M} N={N} mask={masked}: co_attn fwd {f:.2f} ms, bwd (dQ + dK/dV) {b:.2f} ms, pack_mask {pm:.2f} ms | "
          f"torch SDPA fwd {sf:.2f} ms, bwd {sb:.2f} ms")
