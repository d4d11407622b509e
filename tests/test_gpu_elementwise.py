"""GPU: memory-bound kernels against torch fp32 references, through the C ABI."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import err_stats, op_depthwise, op_layernorm, op_lengths

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("d", [64, 176, 256, 512, 1024])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_layernorm(d, out_dtype):
    torch.manual_seed(d)
    rows = 333
    x = torch.randn(rows, d, device="cuda") * 3 + 1.5
    g = torch.randn(d, device="cuda")
    b = torch.randn(d, device="cuda")
    want = F.layer_norm(x, (d,), g, b, 1e-5)
    out = torch.full((rows, d), float("nan"), device="cuda", dtype=out_dtype)
    op_layernorm(x, g, b, out)
    st = err_stats(out.float(), want)
    assert st["nan"] == 0
    assert st["max_abs"] < (2e-5 if out_dtype == torch.float32 else 5e-2) and st["rel_l2"] < (1e-6 if out_dtype == torch.float32 else 4e-3), st


def test_layernorm_in_place_and_masked():
    torch.manual_seed(0)
    B, T, d = 3, 21, 176
    x = torch.randn(B * T, d, device="cuda")
    g, b = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    lens = torch.tensor([21, 9, 0], dtype=torch.int32, device="cuda")
    want = F.layer_norm(x, (d,), g, b, 1e-5)
    valid = (torch.arange(T, device="cuda")[None] < lens[:, None]).reshape(-1, 1)
    y = x.clone()
    op_layernorm(y, g, b, y)  # in place (norm_out of the inner layers)
    assert err_stats(y, want)["max_abs"] < 2e-5
    out = torch.full_like(x, float("nan"))
    op_layernorm(x, g, b, out, lens=lens, frames_per_seq=T)
    assert err_stats(out, want * valid)["max_abs"] < 2e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,d,k", [(2, 50, 176, 31), (3, 130, 512, 31), (1, 7, 64, 31), (2, 64, 256, 9)])
def test_depthwise_bn_swish(dtype, B, T, d, k):
    torch.manual_seed(T)
    x = torch.randn(B, T, d, device="cuda").to(dtype)
    taps = torch.randn(d, k, device="cuda") / k ** 0.5
    bias = torch.randn(d, device="cuda") * 0.1
    want = F.silu(F.conv1d(x.float().transpose(1, 2), taps.unsqueeze(1), bias, padding=(k - 1) // 2, groups=d)).transpose(1, 2)
    out = torch.full((B, T, d), float("nan"), device="cuda", dtype=dtype)
    op_depthwise(x, taps, bias, out)
    st = err_stats(out.float(), want)
    assert st["nan"] == 0 and st["max_abs"] < (2e-5 if dtype == torch.float32 else 3e-2), st


def test_lengths_bit_exact_against_reference_golden():
    z = np.load(os.path.join(GOLDEN, "lengths.npz"))
    lengths = torch.from_numpy(z["lengths"]).cuda()
    for rep in (1, 2, 3):
        got = op_lengths(lengths, lengths.numel(), 0, rep)
        assert got.dtype == torch.int32
        assert np.array_equal(got.cpu().numpy(), z[f"rep{rep}"]), rep
    # length=None: every row is T_full
    got = op_lengths(None, 5, 2001, 2)
    assert got.tolist() == [501] * 5
