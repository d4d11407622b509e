"""GPU: the residual GEMM with the following LayerNorm(s) in its epilogue (gemm_lnc.cu, cfb_op_gemm_ln) against torch
fp32 on the bf16-rounded operands, and against the two-kernel path it replaces (cfb_op_gemm RESID + cfb_op_layernorm).
Reference semantics: conformer_modules.py:98-118 (residual updates), :120 (norm_out) and the next block's input
LayerNorm (:98, :103, :112)."""
import pytest
import torch
import torch.nn.functional as F

from conformer_nemo_b200 import _lib
from gpu_util import err_stats, op_gemm, op_gemm_ln, op_layernorm

pytestmark = pytest.mark.gpu


def make(M, N, K, seed, mean_shift=0.3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn(M, K, generator=g, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, generator=g, device="cuda") * 0.1
    x = torch.randn(M, N, generator=g, device="cuda") * 2 + mean_shift
    ln = [(torch.rand(N, generator=g, device="cuda") + 0.5, torch.randn(N, generator=g, device="cuda") * 0.1)
          for _ in range(2)]
    return A, W, bias, x, ln


# (rows, d_model, K): d = 512 runs on CTA pairs, d <= 256 on single CTAs; 176 = the Small recipe (column tail inside
# a 32-column chunk), 200 / 328 / 456: odd chunk counts per quarter row; row tails; one and many row blocks per CTA
SHAPES = [(128, 512, 512), (300, 512, 2048), (1000, 256, 256), (77, 176, 704), (129, 64, 64), (16000, 512, 512),
          (8000, 512, 2048), (515, 328, 128), (260, 456, 64), (333, 200, 256), (40000, 256, 1024)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("dual", [False, True])
def test_gemm_ln_matches_torch(M, N, K, dual):
    A, W, bias, x, ln = make(M, N, K, 3)
    alpha = 0.5
    v = x + alpha * (A.float() @ W.float().t() + bias)
    y = F.layer_norm(v, (N,), ln[0][0], ln[0][1], 1e-5) if dual else v
    z = F.layer_norm(y, (N,), ln[1][0], ln[1][1], 1e-5)
    xs = x.clone()
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    op_gemm_ln(A, W, bias, alpha, xs, ln[0] if dual else None, ln[1], out)
    s1 = err_stats(xs, y)
    s2 = err_stats(out.float(), z)
    assert s1["nan"] == 0 and s1["max_abs"] < 2e-3, s1          # fp32 stream: only accumulation-order differences
    assert s2["nan"] == 0 and s2["rel_l2"] < 4e-3 and s2["max_abs"] < 5e-2, s2  # one bf16 rounding


@pytest.mark.parametrize("M,N,K", [(16000, 512, 2048), (5000, 256, 256), (700, 176, 176)])
def test_gemm_ln_against_the_two_kernel_path(M, N, K):
    """The stream must be BIT-identical to the reduce-add epilogue's (same arithmetic: x + fma(alpha, acc, alpha b)); the
    normalised operand differs from the stand-alone LayerNorm only by the order in which the row statistics are summed."""
    A, W, bias, x, ln = make(M, N, K, 11)
    x_ref = x.clone()
    op_gemm(True, _lib.EPI_RESID, A, W, bias=bias, out=x_ref, alpha=0.5)
    a_ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    op_layernorm(x_ref, ln[1][0], ln[1][1], a_ref)
    xs = x.clone()
    a = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    op_gemm_ln(A, W, bias, 0.5, xs, None, ln[1], a)
    assert torch.equal(xs, x_ref)
    d = (a.float() - a_ref.float()).abs()
    # bf16 outputs: a last-bit difference in the statistics can flip a rounding (one bf16 ulp at |a| < 8 is 0.03)
    assert d.max().item() <= 0.0625 and (d > 0).float().mean().item() < 5e-3, (d.max().item(), (d > 0).float().mean().item())


def test_gemm_ln_large_row_mean_and_repeatability():
    """Rows whose mean is large against their spread (the single-pass E[x^2] - E[x]^2 form would lose the variance; the
    pairwise (mean, M2) merge does not), and two runs give the same bits."""
    M, N, K = 1000, 512, 512
    A, W, bias, x, ln = make(M, N, K, 5, mean_shift=300.0)
    v = x + (A.float() @ W.float().t() + bias)
    z = F.layer_norm(v, (N,), ln[1][0], ln[1][1], 1e-5)
    outs = []
    for _ in range(2):
        xs = x.clone()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        op_gemm_ln(A, W, bias, 1.0, xs, None, ln[1], out)
        outs.append((xs, out))
    assert err_stats(outs[0][1].float(), z)["rel_l2"] < 4e-3
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_gemm_ln_rows_are_independent():
    """A row's result depends on nothing but the row (what packed == dense rests on): the same rows inside batches of
    different sizes, at different positions of a row block, give the same bits."""
    N, K = 512, 2048
    A, W, bias, x, ln = make(3000, N, K, 7)
    xs = x.clone()
    out = torch.empty(3000, N, device="cuda", dtype=torch.bfloat16)
    op_gemm_ln(A, W, bias, 0.5, xs, ln[0], ln[1], out)
    sel = slice(1001, 1001 + 517)
    A2, x2 = A[sel].contiguous(), x[sel].clone()
    out2 = torch.empty(517, N, device="cuda", dtype=torch.bfloat16)
    op_gemm_ln(A2, W, bias, 0.5, x2, ln[0], ln[1], out2)
    assert torch.equal(x2, xs[sel]) and torch.equal(out2, out[sel])
