"""CPU tests: the log-mel front-end oracle (oracle/frontend_oracle.py) against the committed golden vectors produced
by the unmodified reference FilterbankFeatures (tests/golden/make_golden_frontend.py) and, when /root/reference is
present, against the reference executed live."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as fo
from oracle import reference_loader as rl

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "frontend_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    torch.set_num_threads(1)
    z = load(name)
    y, yl = fo.filterbank_features(z["audio"], z["lengths"], z["window"], z["fb"][0], pad_to=int(z["pad_to"]))
    assert torch.equal(yl, z["seq_len"]) and yl.dtype == torch.int64
    assert y.shape == z["features"].shape
    assert float((y - z["features"]).abs().max()) <= 1e-5


def test_mel_filters_restated_from_librosa_unpinned_are_slaney_unit_area_triangles():
    fb = fo.slaney_mel_filters(16000, 512, 80, 0.0, None)
    assert fb.shape == (80, 257) and fb.dtype == np.float32 and (fb >= 0).all()
    z = load(CASES[0])
    assert np.array_equal(fb, z["fb"][0].numpy())  # what the golden run handed to the reference
    # every filter is a single bump, neighbouring filters overlap, area is 2 / bandwidth * bandwidth / 2 * ... ~ const
    peaks = fb.argmax(1)
    assert (np.diff(peaks) > 0).all()
    freqs = np.linspace(0, 8000, 257)
    area = (fb * (freqs[1] - freqs[0])).sum(1)
    assert np.allclose(area[5:], 1.0, atol=0.12)


def test_seq_len_formula_edges():
    lens = torch.tensor([1, 159, 160, 161, 400, 16000, 320000], dtype=torch.int64)
    assert fo.seq_len_frames(lens, 512, 160).tolist() == [1, 1, 2, 2, 3, 101, 2001]


@pytest.mark.skipif(not rl.reference_available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference():
    torch.set_num_threads(1)
    cls = rl.load_reference_filterbank_class()
    ref = cls(sample_rate=16000, n_window_size=400, n_window_stride=160, n_fft=512, nfilt=80, pad_to=16).eval()
    x, lens = fo.synthetic_waveforms(2, 9000, [9000, 5000], 7)
    with torch.no_grad():
        want, wl = ref(x.clone(), lens)
    got, gl = fo.filterbank_features(x, lens, ref.window, ref.fb[0], pad_to=16)
    assert torch.equal(gl, wl) and float((got - want).abs().max()) <= 1e-6
