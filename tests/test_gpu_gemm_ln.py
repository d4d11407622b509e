"""GPU: the row-complete GEMM with fused residual update + LayerNorm(s) (gemm_ln.cu) against torch fp32 on the
bf16-rounded operands.  Reference semantics: conformer_modules.py:98-120 (residual updates, norm_out) followed by the
next block's input LayerNorm (:98, :103, :112, :116)."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import err_stats, op_gemm_ln

pytestmark = pytest.mark.gpu


def make(M, N, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn(M, K, generator=g, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, generator=g, device="cuda") * 0.1
    resid = torch.randn(M, N, generator=g, device="cuda") * 2 + 0.3
    ln = [(torch.rand(N, generator=g, device="cuda") + 0.5, torch.randn(N, generator=g, device="cuda") * 0.1)
          for _ in range(2)]
    return A, W, bias, resid, ln


SHAPES = [(128, 512, 512), (300, 512, 2048), (1000, 256, 256), (77, 176, 704), (129, 64, 64), (16000, 512, 512)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("mode", ["resid_ln2", "resid_ln1_ln2", "linear_ln2", "resid_ln1_only", "plain"])
def test_gemm_ln_matches_torch(M, N, K, mode):
    A, W, bias, resid, ln = make(M, N, K, 3)
    alpha = 0.5
    use_resid = mode.startswith("resid")
    ln1 = ln[0] if "ln1" in mode else None
    ln2 = ln[1] if "ln2" in mode else None
    v = alpha * (A.float() @ W.float().t() + bias)
    if use_resid:
        v = v + resid
    y = F.layer_norm(v, (N,), ln1[0], ln1[1], 1e-5) if ln1 is not None else v
    z = F.layer_norm(y, (N,), ln2[0], ln2[1], 1e-5) if ln2 is not None else y
    out_f32 = resid.clone() if use_resid else torch.full((M, N), float("nan"), device="cuda")
    out_bf16 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    op_gemm_ln(A, W, bias, alpha, out_f32 if use_resid else None, ln1, ln2, out_f32, out_bf16)  # in place on the stream
    s1 = err_stats(out_f32, y)
    s2 = err_stats(out_bf16.float(), z)
    assert s1["nan"] == 0 and s1["max_abs"] < 2e-3, s1          # fp32 stream: only accumulation-order differences
    assert s2["nan"] == 0 and s2["rel_l2"] < 4e-3 and s2["max_abs"] < 5e-2, s2  # one bf16 rounding


def test_gemm_ln_masks_padded_frames_and_single_outputs():
    M, N, K = 3 * 50, 512, 512
    A, W, bias, resid, ln = make(M, N, K, 5)
    lens = torch.tensor([50, 17, 0], dtype=torch.int32, device="cuda")
    v = resid + (A.float() @ W.float().t() + bias)
    y = F.layer_norm(v, (N,), ln[0][0], ln[0][1], 1e-5)
    keep = (torch.arange(50, device="cuda")[None] < lens[:, None]).reshape(M, 1)
    # fp32 output only (the last layer writing `encoded`)
    out = torch.full((M, N), float("nan"), device="cuda")
    op_gemm_ln(A, W, bias, 1.0, resid, ln[0], None, out, None, lens, 50)
    assert torch.all(out[~keep.expand_as(out)] == 0)
    assert err_stats(out * keep, y * keep)["max_abs"] < 2e-3
    # bf16 output only
    outb = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    op_gemm_ln(A, W, bias, 1.0, resid, ln[0], None, None, outb, lens, 50)
    assert torch.all(outb[~keep.expand_as(outb)] == 0)
    assert err_stats(outb.float() * keep, y * keep)["rel_l2"] < 4e-3
