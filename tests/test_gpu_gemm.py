"""GPU: tcgen05 3xTF32 GEMM (co_gemm_tf32x3) against a float64 reference; accuracy must be
fp32-class (a plain TF32 GEMM would be ~1e-3 relative and fail)."""

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias, residual, scale, shift, relu):
    c = a.double() @ w.double().t()
    if bias is not None:
        c = c + bias.double()
    if residual is not None:
        c = c + residual.double()
    if relu:
        c = c.clamp_min(0)
    if scale is not None:
        c = c * scale.double() + shift.double()
    return c


@pytest.mark.parametrize("M,K,Nout", [(128, 128, 128), (1000, 128, 384), (777, 512, 128), (4096, 128, 640), (130, 128, 512),
                                      (76877, 128, 640), (70001, 128, 128), (80000, 128, 384)])  # last 3: W-stationary path
@pytest.mark.parametrize("epi", ["plain", "bias_relu", "residual_affine"])
def test_gemm_tf32x3_matches_fp64(M, K, Nout, epi):
    from rl4co_b200 import native

    dev = torch.device("cuda:0")
    torch.manual_seed(M + K + Nout)
    a = torch.randn(M, K, device=dev)
    w = torch.randn(Nout, K, device=dev) / K ** 0.5
    bias = residual = scale = shift = None
    relu = False
    if epi == "bias_relu":
        bias, relu = torch.randn(Nout, device=dev), True
    if epi == "residual_affine":
        bias = torch.randn(Nout, device=dev)
        residual = torch.randn(M, Nout, device=dev)
        scale, shift = torch.rand(Nout, device=dev) + 0.5, torch.randn(Nout, device=dev)
    hi, lo = native.split_tf32(w)
    assert torch.equal(hi + lo, w)
    assert ((hi.view(torch.int32) & 0x1FFF) == 0).all()
    out = native.gemm_tf32x3(a, hi, lo, bias=bias, residual=residual, scale=scale, shift=shift, relu=relu)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias, residual, scale, shift, relu)
    err = (out.double() - ref).abs().max().item()
    fp32 = torch.nn.functional.linear(a, w)  # cuBLAS fp32 SIMT for scale
    err32 = (_ref(a, w, None, None, None, None, False) - fp32.double()).abs().max().item()
    mag = ref.abs().max().item()
    assert err <= max(8 * err32, 4e-6 * mag), f"err {err:.3e} vs fp32 {err32:.3e} (|C| {mag:.2f})"


def test_gemm_strided_views_and_column_block_output():
    """A as a row-strided view and C written into a column block of a wider buffer (how the
    encoder / cache use it)."""
    from rl4co_b200 import native

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    big = torch.randn(300, 384, device=dev)
    a = big[:, 128:256]
    w = torch.randn(256, 128, device=dev) * 0.1
    hi, lo = native.split_tf32(w)
    outbuf = torch.zeros(300, 640, device=dev)
    native.gemm_tf32x3(a, hi, lo, out=outbuf[:, 128:384])
    ref = a.double() @ w.double().t()
    torch.testing.assert_close(outbuf[:, 128:384].double(), ref, rtol=1e-5, atol=1e-5)
    assert (outbuf[:, :128] == 0).all() and (outbuf[:, 384:] == 0).all()
