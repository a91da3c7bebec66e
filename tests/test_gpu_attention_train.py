"""co_attn_fwd / co_attn_bwd (training-step attention) against float64 torch attention and its
autograd gradients: rl4co/models/nn/attention.py:110-134 (encoder), :300-314 (masked glimpse)."""
import math

import pytest
import torch

from rl4co_b200 import attention_train as AT


def ref_attention(q, k, v, mask=None):
    B, M, E = q.shape
    H = 8

    def heads(x):
        return x.view(B, x.shape[1], H, E // H).transpose(1, 2)

    s = heads(q) @ heads(k).transpose(-1, -2) / math.sqrt(E // H)
    if mask is not None:
        s = s.masked_fill(~mask[:, None], float("-inf"))
    return (torch.softmax(s, -1) @ heads(v)).transpose(1, 2).reshape(B, M, E)


def test_pack_mask_bits_cpu():
    torch.manual_seed(0)
    m = torch.rand(3, 7, 101) < 0.5
    w = AT.pack_mask(m)
    assert w.shape == (3, 7, 4) and w.dtype == torch.int32
    for n in (0, 1, 31, 32, 63, 64, 100):
        bit = (w[..., n // 32].to(torch.int64) >> (n % 32)) & 1
        assert torch.equal(bit.bool(), m[..., n])
    assert int(((w[..., 3].to(torch.int64) & 0xFFFFFFFF) >> (101 - 96)).sum()) == 0  # padding keys are masked


@pytest.mark.gpu
@pytest.mark.parametrize("B,M,N,masked", [(5, 101, 101, False), (3, 131, 101, True), (4, 20, 20, True), (2, 1, 51, True),
                                          (2, 256, 128, True), (3, 50, 51, False)])
def test_attention_forward_backward_vs_float64(B, M, N, masked):
    dev = torch.device("cuda:0")
    torch.manual_seed(B * 1000 + M)
    q = torch.randn(B, M, 128, device=dev, requires_grad=True)
    cache = torch.randn(B, N, 4 * 128, device=dev, requires_grad=True)   # K / V as column views of a wider tensor
    mask = None
    if masked:
        mask = torch.rand(B, M, N, device=dev) < 0.6
        mask[..., 0] = True  # never a fully masked row
    k, v = cache[..., :128], cache[..., 128:256]
    o = AT.attention(q, k, v, mask)
    g = torch.randn_like(o)
    (o * g).sum().backward()
    q64 = q.detach().double().requires_grad_(True)
    c64 = cache.detach().double().requires_grad_(True)
    o64 = ref_attention(q64, c64[..., :128], c64[..., 128:256], mask)
    (o64 * g.double()).sum().backward()
    torch.testing.assert_close(o.double(), o64, rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(q.grad.double(), q64.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(cache.grad.double(), c64.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_attention_query_chunks_accumulate_key_gradients():
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    B, M, N = 2, 600, 100
    q = torch.randn(B, M, 128, device=dev, requires_grad=True)
    k = torch.randn(B, N, 128, device=dev, requires_grad=True)
    v = torch.randn(B, N, 128, device=dev, requires_grad=True)
    mask = torch.rand(B, M, N, device=dev) < 0.7
    mask[..., 3] = True
    o = AT.attention(q, k, v, mask)
    o.square().sum().backward()
    q64, k64, v64 = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    o64 = ref_attention(q64, k64, v64, mask)
    o64.square().sum().backward()
    torch.testing.assert_close(o.double(), o64, rtol=1e-5, atol=2e-6)
    for a, b in ((q, q64), (k, k64), (v, v64)):
        torch.testing.assert_close(a.grad.double(), b.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_self_attention_packed_matches_sdpa_and_gradients():
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    B, N = 6, 101
    qkv = torch.randn(B, N, 384, device=dev, requires_grad=True)
    o = AT.self_attention_packed(qkv)
    g = torch.randn_like(o)
    (o * g).sum().backward()
    x64 = qkv.detach().double().requires_grad_(True)
    o64 = ref_attention(x64[..., :128], x64[..., 128:256], x64[..., 256:])
    (o64 * g.double()).sum().backward()
    torch.testing.assert_close(o.double(), o64, rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(qkv.grad.double(), x64.grad, rtol=1e-4, atol=1e-5)
