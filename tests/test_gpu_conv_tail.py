"""GPU: the fused tail of the convolution module (depth-wise conv + folded BatchNorm + Swish + pointwise_conv2 +
residual add, csrc/conv_tail.cu) against a torch fp32 reference of conformer_modules.py:168-180 and against the two
stand-alone kernels it replaces, through the C ABI."""
import pytest
import torch
import torch.nn.functional as F

from conformer_nemo_b200 import _lib
from gpu_util import err_stats, op_depthwise, op_dw_pw2, op_gemm

pytestmark = pytest.mark.gpu


def make_case(B, T, d, k, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    g = torch.randn(B, T, d, device="cuda", generator=gen).to(torch.bfloat16)
    taps = torch.randn(d, k, device="cuda", generator=gen) / k ** 0.5
    bias = torch.randn(d, device="cuda", generator=gen) * 0.1
    W2 = (torch.randn(d, d, device="cuda", generator=gen) / d ** 0.5).to(torch.bfloat16)
    bias2 = torch.randn(d, device="cuda", generator=gen) * 0.1
    x = torch.randn(B, T, d, device="cuda", generator=gen)
    return g, taps, bias, W2, bias2, x


@pytest.mark.parametrize("B,T,d,k", [(2, 128, 512, 31), (3, 130, 512, 31), (1, 7, 64, 31), (2, 100, 256, 31),
                                     (4, 500, 512, 31), (2, 64, 256, 9), (1, 300, 192, 15), (5, 257, 384, 31),
                                     (1, 1, 128, 31)])
def test_fused_tail_matches_reference_and_unfused(B, T, d, k):
    g, taps, bias, W2, bias2, x = make_case(B, T, d, k, seed=T * 7 + d)
    # torch fp32 reference on the bf16-rounded operands
    c = F.silu(F.conv1d(g.float().transpose(1, 2), taps.unsqueeze(1), bias, padding=(k - 1) // 2, groups=d)).transpose(1, 2)
    want = x + c @ W2.float().t() + bias2
    got = op_dw_pw2(g, taps, bias, W2, bias2, x.clone())
    st = err_stats(got, want)
    assert st["nan"] == 0 and st["max_abs"] < 3e-2 and st["rel_l2"] < 3e-3, st
    # the two kernels it replaces: same arithmetic in the same order -> identical bits
    cb = torch.empty_like(g)
    op_depthwise(g, taps, bias, cb)
    x2 = x.clone().reshape(B * T, d)
    op_gemm(True, _lib.EPI_RESID, cb.reshape(B * T, d), W2, bias=bias2, out=x2, alpha=1.0)
    assert torch.equal(got.reshape(B * T, d), x2), err_stats(got.reshape(B * T, d), x2)


def test_fused_tail_does_not_touch_neighbouring_sequences():
    """Frames outside [0, T) of a sequence are the conv's zero padding, not the neighbouring sequence's frames."""
    B, T, d, k = 3, 40, 128, 31
    g, taps, bias, W2, bias2, x = make_case(B, T, d, k, seed=5)
    full = op_dw_pw2(g, taps, bias, W2, bias2, x.clone())
    for b in range(B):
        one = op_dw_pw2(g[b:b + 1].contiguous(), taps, bias, W2, bias2, x[b:b + 1].clone())
        assert torch.equal(one[0], full[b])


def test_fused_tail_is_repeatable():
    """Twenty launches on the same operands give the same bits (no race between the SIMT producers, the TMA loads and
    the MMA issuer; one reduce-add per output element)."""
    B, T, d, k = 8, 500, 512, 31
    g, taps, bias, W2, bias2, x = make_case(B, T, d, k, seed=11)
    first = op_dw_pw2(g, taps, bias, W2, bias2, x.clone())
    for _ in range(19):
        assert torch.equal(op_dw_pw2(g, taps, bias, W2, bias2, x.clone()), first)
