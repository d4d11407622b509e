"""CPU tests: the oracle (oracle/conformer_oracle.py) against the committed golden vectors produced by the
unmodified reference (tests/golden/make_golden.py), and -- when /root/reference is present -- against the
reference executed live."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import conformer_oracle as oc
from oracle import reference_loader as rl

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "lengths" not in p and not os.path.basename(p).startswith(("ctc_head_", "ctc_collapse", "frontend_", "rnnt_")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    cfg = oc.EncoderConfig(**meta["config"])
    stored = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    sd = stored if stored else oc.random_state_dict(cfg, meta["weight_seed"])
    return z, meta, cfg, sd


@pytest.mark.parametrize("name", CASES)
def test_weight_fixture_checksums(name):
    z, meta, cfg, sd = load_case(name)
    assert set(sd) == set(meta["checksums"])
    for k, v in sd.items():
        assert float(v.double().sum()) == pytest.approx(meta["checksums"][k], rel=1e-12, abs=1e-12), k
        assert tuple(v.shape) == oc.expected_state_shapes(cfg)[k]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    torch.set_num_threads(1)
    z, meta, cfg, sd = load_case(name)
    x = torch.from_numpy(z["audio_signal"])
    length = torch.from_numpy(z["length"]) if bool(z["has_length"]) else None
    y, ylen = oc.encoder_forward(sd, cfg, x, length)
    assert ylen.dtype == torch.int32
    assert np.array_equal(ylen.numpy(), z["encoded_len"])
    assert tuple(y.shape) == z["encoded"].shape
    # same algorithm, same torch kernels: agreement to fp32 round-off, on every frame (padded ones included)
    np.testing.assert_allclose(y.numpy(), z["encoded"], rtol=0, atol=2e-5)


def test_lengths_golden():
    z = np.load(os.path.join(GOLDEN, "lengths.npz"))
    lengths = torch.from_numpy(z["lengths"])
    for rep in (1, 2, 3):
        got = oc.subsampled_lengths(lengths, rep)
        assert got.dtype == torch.int32
        assert np.array_equal(got.numpy(), z[f"rep{rep}"])
    # SURVEY section 4.4 known-answer vector
    assert oc.subsampled_lengths(torch.tensor([1000, 900, 640, 333]), 2).tolist() == [250, 225, 160, 84]


def test_rel_shift_gather_is_the_pad_view_trick():
    """multi_head_attention.py:159-170: pad one zero column, view (b,h,2T,T), drop the first row, view back."""
    torch.manual_seed(0)
    b, h, t = 2, 3, 7
    x = torch.randn(b, h, t, 2 * t - 1)
    padded = torch.nn.functional.pad(x, (1, 0)).view(b, h, -1, t)[:, :, 1:].view(b, h, t, 2 * t - 1)[..., :t]
    assert torch.equal(oc.rel_shift_gather(x), padded)


def test_pos_table_layout():
    t, d = 5, 8
    tab = oc.rel_pos_table(t, d)
    assert tab.shape == (2 * t - 1, d)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(np.log(10000.0) / d))
    for k in range(2 * t - 1):
        pos = float(t - 1 - k)
        assert torch.allclose(tab[k, 0::2], torch.sin(pos * div))
        assert torch.allclose(tab[k, 1::2], torch.cos(pos * div))


def test_fully_masked_query_rows_give_bias_only():
    """a padded query row has zero context, so the attention block returns linear_out.bias."""
    cfg = oc.EncoderConfig(feat_in=80, n_layers=1, d_model=64, n_heads=4)
    sd = oc.random_state_dict(cfg, 7)
    x = torch.randn(2, 9, 64)
    valid = torch.arange(9).unsqueeze(0) < torch.tensor([9, 4]).unsqueeze(1)
    stages = {}
    out = oc._rel_pos_attention(sd, "layers.0.self_attn", x, oc.rel_pos_table(9, 64), valid, 4, stages)
    assert torch.all(stages["ctx"][1, 4:] == 0)
    assert torch.allclose(out[1, 4:], sd["layers.0.self_attn.linear_out.bias"].expand(5, -1))


@pytest.mark.skipif(not rl.reference_available(), reason="reference tree not mounted (GPU box)")
def test_oracle_matches_live_reference_mixed_lengths():
    torch.set_num_threads(1)
    cfg = oc.EncoderConfig(feat_in=80, n_layers=2, d_model=96, n_heads=4)
    sd = oc.random_state_dict(cfg, 11)
    enc = rl.build_reference_encoder(cfg, sd)
    x, length = oc.synthetic_batch(4, 80, 131, [131, 77, 130, 2], seed=99)
    with torch.no_grad():
        yr, lr = enc(audio_signal=x, length=length)
    yo, lo = oc.encoder_forward(sd, cfg, x, length)
    assert torch.equal(lr, lo)
    np.testing.assert_allclose(yo.numpy(), yr.numpy(), rtol=0, atol=2e-5)
