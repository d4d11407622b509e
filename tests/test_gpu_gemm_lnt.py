"""GPU: GEMM with a LayerNorm prologue whose operand lives in tensor memory (gemm_lnt.cu) == stand-alone LayerNorm
kernel + tcgen05 GEMM, bit for bit, and close to torch fp32 (conformer_modules.py:98-120)."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import err_stats, op_gemm, op_gemm_lna, op_layernorm

pytestmark = pytest.mark.gpu
EPI = dict(LINEAR=0, SWISH=1, QKV=4, GLU=5)


def make(M, d, N, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(M, d, generator=g, device="cuda") * 1.7 + 0.2
    W = (torch.randn(N, d, generator=g, device="cuda") / d ** 0.5).bfloat16()
    bias = torch.randn(N, generator=g, device="cuda") * 0.1
    ln = [(torch.rand(d, generator=g, device="cuda") + 0.5, torch.randn(d, generator=g, device="cuda") * 0.1)
          for _ in range(2)]
    return x, W, bias, ln


@pytest.mark.parametrize("M,d,N,epi", [
    (128, 512, 2048, "SWISH"), (300, 512, 2048, "SWISH"), (16000, 512, 2048, "SWISH"), (1000, 256, 1024, "SWISH"),
    (77, 64, 256, "SWISH"), (500, 512, 1024, "GLU"), (333, 256, 512, "GLU"), (640, 512, 1536, "QKV"),
    (250, 192, 768, "QKV"), (200, 512, 256, "LINEAR"), (25600, 256, 1024, "SWISH"), (19000, 512, 1536, "QKV"),
])
@pytest.mark.parametrize("dual", [False, True])
def test_gemm_lnt_equals_layernorm_then_gemm(M, d, N, epi, dual):
    x, W, bias, ln = make(M, d, N, 11)
    frames = 50
    lens = torch.randint(0, frames + 1, ((M + frames - 1) // frames,), device="cuda", dtype=torch.int32) if epi == "GLU" else None
    qkv_dp = N // 3 if epi == "QKV" else 0
    bias2 = bias[:qkv_dp].clone() * 0.5 if epi == "QKV" else None
    ncols = {"QKV": N + qkv_dp, "GLU": N // 2}.get(epi, N)
    y = x
    if dual:
        y = torch.empty_like(x)
        op_layernorm(x, ln[0][0], ln[0][1], y)
    a = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
    op_layernorm(y, ln[1][0], ln[1][1], a)
    want = torch.zeros(M, ncols, device="cuda", dtype=torch.bfloat16)
    op_gemm(True, EPI[epi], a, W, bias, bias2, want, 1.0, lens, frames, qkv_dp)
    got = torch.full((M, ncols), float("nan"), device="cuda", dtype=torch.bfloat16)
    x_out = torch.full_like(x, float("nan")) if dual else None
    op_gemm_lna(EPI[epi], x, W, ln[1], got, bias, bias2, ln[0] if dual else None, x_out, lens, frames, qkv_dp, tmem=True)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16)), err_stats(got.float(), want.float())
    if dual:
        assert torch.equal(x_out, y)
    if epi in ("LINEAR", "SWISH"):
        z = F.layer_norm(F.layer_norm(x, (d,), *ln[0], 1e-5) if dual else x, (d,), *ln[1], 1e-5)
        ref = z.bfloat16().float() @ W.float().t() + bias
        if epi == "SWISH":
            ref = F.silu(ref)
        assert err_stats(got.float(), ref)["rel_l2"] < 6e-3


def test_gemm_lnt_in_place_norm_out():
    """x_out may alias x (norm_out rewrites the residual stream in place)."""
    M, d, N = 700, 512, 2048
    x, W, bias, ln = make(M, d, N, 3)
    y = torch.empty_like(x)
    op_layernorm(x, ln[0][0], ln[0][1], y)
    a = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
    op_layernorm(y, ln[1][0], ln[1][1], a)
    want = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    op_gemm(True, EPI["SWISH"], a, W, bias, None, want, 1.0, None, 1, 0)
    xin = x.clone()
    got = torch.empty_like(want)
    op_gemm_lna(EPI["SWISH"], xin, W, ln[1], got, bias, None, ln[0], xin, None, 1, 0, tmem=True)
    assert torch.equal(xin, y) and torch.equal(got.view(torch.int16), want.view(torch.int16))
