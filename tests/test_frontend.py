"""Log-mel front-end: host-side mirror of AudioToMelSpectrogramPreprocessor (CPU tests) and the
CUDA kernels behind cfb_op_logmel against the golden vectors of the unmodified reference and against the CPU oracle
(GPU tests).  Tolerance: fp32 FFT / reduction order differences only -- max-abs 2e-3 on unit-variance features,
rel-L2 2e-4; processed_length bit-exact (int64)."""
import glob
import os

import numpy as np
import pytest
import torch

import conformer_nemo_b200 as cn
from oracle import frontend_oracle as fo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "frontend_*.npz")))


def recipe(**kw):
    args = dict(sample_rate=16000, window_size=0.025, window_stride=0.01, window="hann", normalize="per_feature",
                n_fft=512, features=80, dither=1e-5, pad_to=0)  # conformer_ctc_bpe.yaml model.preprocessor
    args.update(kw)
    return cn.AudioToMelSpectrogramPreprocessor(**args)


def test_state_dict_and_buffers_match_the_reference_layout():
    p = recipe()
    sd = p.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {"featurizer.window": (400,), "featurizer.fb": (1, 80, 257)}
    z = np.load(os.path.join(GOLDEN, CASES[0] + ".npz"))
    assert np.array_equal(sd["featurizer.window"].numpy(), z["window"])  # torch.hann_window(400, periodic=False)
    assert np.array_equal(sd["featurizer.fb"].numpy(), z["fb"])
    assert p.get_seq_len(torch.tensor([1, 160, 16000, 320000])).tolist() == [1, 2, 101, 2001]
    assert p.filter_banks is p.featurizer.fb


def test_constructor_errors_follow_the_reference():
    with pytest.raises(ValueError):
        recipe(n_window_size=400)  # both window_size and n_window_size (audio_preprocessing.py:238-239)
    with pytest.raises(ValueError):
        recipe(log_zero_guard_type="floor")
    for kw in (dict(normalize="all_features"), dict(frame_splicing=3), dict(exact_pad=True), dict(n_fft=1024),
               dict(mag_power=1.0), dict(log=False), dict(pad_to="max")):
        with pytest.raises(NotImplementedError):
            recipe(**kw)
    with pytest.raises(NotImplementedError):
        recipe().train()
    with pytest.raises(RuntimeError):
        recipe()(input_signal=torch.zeros(1, 1000), length=torch.tensor([1000]))  # no CPU fallback


def _stats(got, want):
    d = (got.double() - want.double()).abs()
    return dict(max_abs=float(d.max()), rel_l2=float(d.norm() / want.double().norm()), nan=int(torch.isnan(got).sum()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    p = recipe(pad_to=int(z["pad_to"])).cuda()
    y, yl = p(input_signal=torch.from_numpy(z["audio"]).cuda(), length=torch.from_numpy(z["lengths"]).cuda())
    torch.cuda.synchronize()
    assert yl.dtype == torch.int64 and np.array_equal(yl.cpu().numpy(), z["seq_len"])
    assert tuple(y.shape) == z["features"].shape
    st = _stats(y.cpu(), torch.from_numpy(z["features"]))
    assert st["nan"] == 0 and st["max_abs"] <= 2e-3 and st["rel_l2"] <= 2e-4, st


@pytest.mark.gpu
@pytest.mark.parametrize("B,L,lens,pad_to", [(4, 80000, [80000, 51234, 16000, 401], 16), (1, 320000, [320000], 0),
                                              (33, 4000, None, 16)])
def test_kernel_matches_oracle(B, L, lens, pad_to):
    if lens is None:
        lens = [int(v) for v in torch.randint(400, L + 1, (B,), generator=torch.Generator().manual_seed(B))]
    x, lengths = fo.synthetic_waveforms(B, L, lens, seed=L % 97)
    p = recipe(pad_to=pad_to)
    want, wl = fo.filterbank_features(x, lengths, p.featurizer.window, p.featurizer.fb[0], pad_to=pad_to)
    p = p.cuda()
    y, yl = p(input_signal=x.cuda(), length=lengths.cuda())
    torch.cuda.synchronize()
    assert torch.equal(yl.cpu(), wl) and y.shape == want.shape
    st = _stats(y.cpu(), want)
    assert st["nan"] == 0 and st["max_abs"] <= 2e-3 and st["rel_l2"] <= 2e-4, st
    for b, n in enumerate(wl.tolist()):
        assert torch.all(y[b, :, n:] == 0)  # frames past processed_length are exactly zero (features.py:436-443)


@pytest.mark.gpu
def test_one_frame_utterance_raises_like_the_reference():
    p = recipe().cuda()
    x = torch.zeros(2, 2000, device="cuda")
    x[0] = torch.randn(2000, device="cuda")
    with pytest.raises(ValueError):
        p(input_signal=x, length=torch.tensor([2000, 100], device="cuda"))  # floor(100 / 160) + 1 = 1 frame
    y, yl = p(input_signal=x, length=torch.tensor([2000, 100], device="cuda"), check_lengths=False)
    assert yl.tolist() == [13, 1]


@pytest.mark.gpu
def test_checkpoint_buffers_are_used():
    """featurizer.fb / featurizer.window loaded from a state_dict replace the generated ones."""
    p = recipe().cuda()
    x, lengths = fo.synthetic_waveforms(2, 6000, [6000, 3000], seed=3)
    sd = {k: v.clone() for k, v in p.state_dict().items()}
    sd["featurizer.fb"] = sd["featurizer.fb"] * 2.0          # log(2 mel) = log(mel) + log 2: the normalisation removes it
    sd["featurizer.window"] = torch.hamming_window(400, periodic=False).cuda()
    y0, _ = p(input_signal=x.cuda(), length=lengths.cuda())
    p.load_state_dict(sd)
    y1, _ = p(input_signal=x.cuda(), length=lengths.cuda())
    want, _ = fo.filterbank_features(x, lengths, sd["featurizer.window"].cpu(), sd["featurizer.fb"][0].cpu(), pad_to=0)
    assert _stats(y1.cpu(), want)["max_abs"] <= 2e-3
    assert float((y1 - y0).abs().max()) > 1e-2  # the window did change the features


@pytest.mark.gpu
def test_waveform_to_encoder_chain_against_the_oracles():
    """waveform -> log-mel kernels -> encoder kernels on the GPU against frontend oracle -> encoder oracle on the CPU."""
    from oracle import conformer_oracle as oc

    x, lengths = fo.synthetic_waveforms(3, 32000, [32000, 20000, 9000], seed=5)
    pre = recipe()
    feats, flen = fo.filterbank_features(x, lengths, pre.featurizer.window, pre.featurizer.fb[0], pad_to=0)
    cfg = oc.EncoderConfig(feat_in=80, n_layers=2, d_model=256, n_heads=4)
    sd = oc.random_state_dict(cfg, 4)
    want, want_len = oc.encoder_forward(sd, cfg, feats, flen)
    enc = cn.ConformerEncoder(feat_in=80, n_layers=2, d_model=256, n_heads=4)
    enc.load_state_dict(sd, strict=False)
    enc = enc.cuda().eval()
    f_gpu, l_gpu = pre.cuda()(input_signal=x.cuda(), length=lengths.cuda())
    y, ylen = enc(audio_signal=f_gpu, length=l_gpu)
    torch.cuda.synchronize()
    assert torch.equal(ylen.cpu(), want_len)
    valid = (torch.arange(want.shape[2])[None] < want_len[:, None].long())[:, None, :].expand_as(want)
    g, w = y.cpu().double()[valid], want.double()[valid]
    assert float((g - w).norm() / w.norm()) <= 1e-2 and float((g - w).abs().max()) <= 5e-2
