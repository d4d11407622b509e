"""GPU: tcgen05 GEMM (+ every fused epilogue) against torch fp32 and against the CUDA-core validation kernel,
called through the C ABI (cfb_op_gemm)."""
import pytest
import torch

from conformer_nemo_b200 import _lib
from gpu_util import describe_mismatch, err_stats, op_gemm

pytestmark = pytest.mark.gpu


def _mk(M, N, K, seed=0, ints=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if ints:  # small integers: every product and partial sum is exact in bf16 x bf16 -> fp32
        A = torch.randint(-3, 4, (M, K), generator=g, device="cuda").float()
        W = torch.randint(-3, 4, (N, K), generator=g, device="cuda").float()
    else:
        A = torch.randn(M, K, generator=g, device="cuda")
        W = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
    return A, W


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (256, 256, 128), (300, 192, 176), (77, 64, 64),
                                   (1000, 1024, 512), (640, 2048, 512), (130, 704, 176)])
def test_tc_linear_exact_on_integers(M, N, K):
    A, W = _mk(M, N, K, ints=True)
    want = A @ W.t()
    out = torch.full((M, N), float("nan"), device="cuda")
    op_gemm(True, _lib.EPI_LINEAR, A.bfloat16(), W.bfloat16(), out=out)
    st = err_stats(out, want)
    assert st["nan"] == 0 and st["max_abs"] == 0.0, f"{st}{describe_mismatch(out, want, 0.0)}"


@pytest.mark.parametrize("use_tc", [False, True])
@pytest.mark.parametrize("M,N,K", [(300, 192, 176), (513, 512, 512)])
def test_linear_bias(use_tc, M, N, K):
    A, W = _mk(M, N, K, seed=1)
    bias = torch.randn(N, device="cuda")
    if use_tc:
        Ab, Wb = A.bfloat16(), W.bfloat16()
        want = Ab.float() @ Wb.float().t() + bias
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        op_gemm(True, _lib.EPI_LINEAR, Ab, Wb, bias=bias, out=out)
        st = err_stats(out.float(), want)
        assert st["nan"] == 0 and st["rel_l2"] < 4e-3, st
    else:
        want = A @ W.t() + bias
        out = torch.zeros(M, N, device="cuda")
        op_gemm(False, _lib.EPI_LINEAR, A, W, bias=bias, out=out)
        st = err_stats(out, want)
        assert st["nan"] == 0 and st["max_abs"] < 1e-4, st


@pytest.mark.parametrize("use_tc", [False, True])
def test_swish_relu(use_tc):
    M, N, K = 260, 704, 176
    A, W = _mk(M, N, K, seed=2)
    bias = torch.randn(N, device="cuda") * 0.5
    dt = torch.bfloat16 if use_tc else torch.float32
    Ai, Wi = A.to(dt), W.to(dt)
    z = Ai.float() @ Wi.float().t() + bias
    for epi, fn in ((_lib.EPI_SWISH, torch.nn.functional.silu), (_lib.EPI_RELU, torch.relu)):
        out = torch.zeros(M, N, device="cuda", dtype=dt)
        op_gemm(use_tc, epi, Ai, Wi, bias=bias, out=out)
        st = err_stats(out.float(), fn(z))
        assert st["nan"] == 0 and st["rel_l2"] < (5e-3 if use_tc else 1e-5), (epi, st)


@pytest.mark.parametrize("use_tc", [False, True])
def test_residual(use_tc):
    M, N, K = 200, 176, 704
    A, W = _mk(M, N, K, seed=3)
    bias = torch.randn(N, device="cuda")
    dt = torch.bfloat16 if use_tc else torch.float32
    Ai, Wi = A.to(dt), W.to(dt)
    resid = torch.randn(M, N, device="cuda")
    want = resid + 0.5 * (Ai.float() @ Wi.float().t() + bias)
    out = resid.clone()
    op_gemm(use_tc, _lib.EPI_RESID, Ai, Wi, bias=bias, out=out, alpha=0.5)
    st = err_stats(out, want)
    assert st["nan"] == 0 and st["max_abs"] < (2e-4 if use_tc else 1e-4), st


@pytest.mark.parametrize("use_tc", [False, True])
def test_qkv_epilogue(use_tc):
    M, K, Dp = 150, 176, 256
    N = 3 * Dp
    A, W = _mk(M, N, K, seed=4)
    bias = torch.randn(N, device="cuda")
    bias2 = torch.randn(Dp, device="cuda")
    dt = torch.bfloat16 if use_tc else torch.float32
    Ai, Wi = A.to(dt), W.to(dt)
    z = Ai.float() @ Wi.float().t()
    want = torch.cat([z[:, :Dp] + bias[:Dp], z[:, :Dp] + bias2, z[:, Dp:] + bias[Dp:]], dim=1)
    out = torch.zeros(M, 4 * Dp, device="cuda", dtype=dt)
    op_gemm(use_tc, _lib.EPI_QKV, Ai, Wi, bias=bias, bias2=bias2, out=out, qkv_dp=Dp)
    st = err_stats(out.float(), want)
    assert st["nan"] == 0 and st["rel_l2"] < (4e-3 if use_tc else 1e-5), st


@pytest.mark.parametrize("use_tc", [False, True])
def test_glu_mask_epilogue(use_tc):
    B, T, d, K = 3, 50, 176, 176
    M, N = B * T, 2 * d
    A, W = _mk(M, N, K, seed=5)
    bias = torch.randn(N, device="cuda")
    lens = torch.tensor([50, 31, 0], dtype=torch.int32, device="cuda")
    dt = torch.bfloat16 if use_tc else torch.float32
    Ai, Wi = A.to(dt), W.to(dt)
    z = Ai.float() @ Wi.float().t() + bias  # reference channel order: [a (d) | gate (d)]
    want = z[:, :d] * torch.sigmoid(z[:, d:])
    valid = (torch.arange(T, device="cuda")[None, :] < lens[:, None]).reshape(M, 1)
    want = want * valid
    # interleave rows the way cfb_finalize_weights does: groups of 32 accumulator columns = [16 a | 16 gate]
    cols = torch.arange(N, device="cuda")
    grp, j = cols // 32, cols % 32
    src = torch.where(j < 16, grp * 16 + j, d + grp * 16 + (j - 16))
    out = torch.full((M, d), float("nan"), device="cuda", dtype=dt)
    op_gemm(use_tc, _lib.EPI_GLU, Ai, Wi[src].contiguous(), bias=bias[src].contiguous(), out=out, lens=lens,
            frames_per_seq=T)
    st = err_stats(out.float(), want)
    assert st["nan"] == 0 and st["rel_l2"] < (5e-3 if use_tc else 1e-5), st


def test_tc_matches_simt_on_bf16_inputs():
    M, N, K = 384, 512, 2048
    A, W = _mk(M, N, K, seed=6)
    Ab, Wb = A.bfloat16(), W.bfloat16()
    o_tc = torch.zeros(M, N, device="cuda")
    o_ref = torch.zeros(M, N, device="cuda")
    op_gemm(True, _lib.EPI_LINEAR, Ab, Wb, out=o_tc)
    op_gemm(False, _lib.EPI_LINEAR, Ab.float(), Wb.float(), out=o_ref)
    st = err_stats(o_tc, o_ref)
    assert st["nan"] == 0 and st["max_abs"] < 1e-4, st


@pytest.mark.parametrize("epi", ["LINEAR", "SWISH", "RESID", "QKV", "GLU"])
@pytest.mark.parametrize("M,N,K", [(129, 256, 64), (300, 512, 176), (1000, 1536, 512), (16000, 2048, 512), (257, 768, 2048)])
def test_cta_pair_kernel_is_bit_identical_to_the_single_cta_kernel(epi, M, N, K, monkeypatch):
    """gemm_tc2.cu (cta_group::2, 256 x 256 tiles per CTA pair) against gemm_tc.cu on the same operands: same
    k-block order and the same epilogue arithmetic -> identical bits, including M tails where the second CTA of the
    last pair owns no valid row.  CFB_GEMM_2CTA forces either kernel (the default picks by shape)."""
    A, W = _mk(M, N, K, seed=5)
    A, W = A.bfloat16(), W.bfloat16()
    bias = torch.randn(N, device="cuda") * 0.3
    frames = 50
    kw = {}
    if epi == "QKV":
        if N % 3:
            pytest.skip("QKV needs N = 3 * Dp")
        dp = N // 3
        kw = dict(bias2=torch.randn(dp, device="cuda") * 0.3, qkv_dp=dp)
        ncols, dt = N + dp, torch.bfloat16
    elif epi == "GLU":
        kw = dict(lens=torch.randint(0, frames + 1, ((M + frames - 1) // frames,), device="cuda", dtype=torch.int32),
                  frames_per_seq=frames)
        ncols, dt = N // 2, torch.bfloat16
    elif epi == "RESID":
        kw = dict(alpha=0.5)
        ncols, dt = N, torch.float32
    else:
        ncols, dt = N, torch.bfloat16
    outs = []
    for mode in ("0", "1", "1"):  # the pair kernel twice: run-to-run repeatability
        monkeypatch.setenv("CFB_GEMM_2CTA", mode)
        gen = torch.Generator(device="cuda").manual_seed(9)
        out = torch.randn(M, ncols, device="cuda", generator=gen).to(dt) if epi == "RESID" else torch.full(
            (M, ncols), float("nan"), device="cuda", dtype=dt)
        op_gemm(True, getattr(_lib, "EPI_" + epi), A, W, bias=bias, out=out, **kw)
        outs.append(out)
    assert int(torch.isnan(outs[0].float()).sum()) == 0
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2]), err_stats(outs[1].float(), outs[0].float())
