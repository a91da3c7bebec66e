"""Time co_gemm_tf32x3 against cuBLAS fp32 (torch) on the path's GEMM shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl4co_b200 import native

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536 * 100
for K, Nout in ((128, 640), (128, 384), (128, 128), (128, 512), (512, 128)):
    a = torch.randn(M, K, device=dev)
    w = torch.randn(Nout, K, device=dev) / K ** 0.5
    hi, lo = native.split_tf32(w)
    out = torch.empty(M, Nout, device=dev)
    def t(fn, n=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_ours = t(lambda: native.gemm_tf32x3(a, hi, lo, out=out))
    t_cublas = t(lambda: torch.nn.functional.linear(a, w, out=None))
    flops = 2.0 * M * K * Nout
    gb = (M * K + M * Nout) * 4 / 1e9
    print(f"M={M} K={K} N={Nout}: tf32x3 {t_ours:7.2f} ms ({flops / t_ours / 1e9:7.1f} TFLOP/s-equiv, {gb / t_ours * 1e3:6.0f} GB/s)  "
          f"cuBLAS fp32 {t_cublas:7.2f} ms  speedup {t_cublas / t_ours:4.2f}x")
    del a, out
