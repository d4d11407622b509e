"""Generates tests/golden/frontend_*.npz by running the UNMODIFIED reference FilterbankFeatures
(parts/preprocessing/features.py, loaded through oracle/reference_loader.py with the stand-ins described there) on
seeded synthetic waveforms.  Build container only:

    python tests/golden/make_golden_frontend.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.frontend_oracle import synthetic_waveforms  # noqa: E402
from oracle.reference_loader import load_reference_filterbank_class  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {  # name: (batch, samples, lengths, pad_to, seed)
    "frontend_mixed": (3, 20000, [20000, 13337, 4000], 0, 0),
    "frontend_pad16": (2, 7777, [7777, 1000], 16, 1),
    "frontend_tiny": (1, 800, [800], 0, 2),
}


def main():
    torch.set_num_threads(1)
    cls = load_reference_filterbank_class()
    for name, (b, n, lens, pad_to, seed) in CASES.items():
        ref = cls(sample_rate=16000, n_window_size=400, n_window_stride=160, window="hann", normalize="per_feature",
                  n_fft=512, nfilt=80, dither=1e-5, pad_to=pad_to).eval()
        x, lengths = synthetic_waveforms(b, n, lens, seed)
        with torch.no_grad():
            y, yl = ref(x.clone(), lengths)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), audio=x.numpy().astype(np.float16 if False else np.float32),
                            lengths=lengths.numpy(), features=y.numpy(), seq_len=yl.numpy(), pad_to=np.array(pad_to),
                            window=ref.window.numpy(), fb=ref.fb.numpy())
        print(name, tuple(y.shape), yl.tolist(), float(y.std()))


if __name__ == "__main__":
    main()
