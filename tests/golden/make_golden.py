"""Generates tests/golden/*.npz by running the UNMODIFIED reference encoder (imported from /root/reference through
oracle/reference_loader.py) on seeded weights and inputs.  Run in the build container only:

    python tests/golden/make_golden.py

Every case stores the input, the lengths, the reference's (encoded, encoded_len) and a per-tensor float64 checksum of
the weights; ``tiny_d64`` additionally stores the full state_dict so the fixture does not depend on torch's RNG stream.
For the other cases the weights are re-created by ``oracle.conformer_oracle.random_state_dict(cfg, seed)`` and checked
against the stored checksums.  ``lengths.npz`` pins ``calc_length`` (subsampling.py:272-282) on edge values.
"""
import dataclasses
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.conformer_oracle import EncoderConfig, random_state_dict, synthetic_batch  # noqa: E402
from oracle.reference_loader import build_reference_encoder  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (config kwargs, weight seed, B, T, lengths, store_weights)
    "tiny_d64": (dict(feat_in=80, n_layers=2, d_model=64, n_heads=4), 0, 3, 83, [83, 60, 17], True),
    "small_d176": (dict(feat_in=80, n_layers=2, d_model=176, n_heads=4), 1, 2, 200, [200, 131], False),
    "char_d256": (dict(feat_in=80, n_layers=1, d_model=256, n_heads=8), 2, 2, 160, [160, 97], False),
    "large_d512": (dict(feat_in=80, n_layers=1, d_model=512, n_heads=8), 3, 2, 240, [240, 100], False),
    "featout_d64": (dict(feat_in=80, n_layers=1, d_model=64, n_heads=4, feat_out=48), 4, 2, 64, [64, 33], False),
    "nolen_d64": (dict(feat_in=80, n_layers=1, d_model=64, n_heads=4), 5, 1, 45, None, False),
    "noxscale_d64": (dict(feat_in=80, n_layers=1, d_model=64, n_heads=4, xscaling=False), 6, 2, 50, [50, 1], False),
    # full depth of the Large recipes (conformer_transducer_bpe.yaml:110), small B / T: pins the oracle's drift over 17 layers
    "large17_d512": (dict(feat_in=80, n_layers=17, d_model=512, n_heads=8), 7, 2, 300, [300, 173], False),
    # constructor variants no other case reaches: shorter depth-wise kernels (taps centred in the 31-tap window), feed-forward
    # expansion factors other than 4, and the tied (encoder-level) pos_bias_u / pos_bias_v pair of untie_biases=False
    "k15_ff2_d64": (dict(feat_in=80, n_layers=2, d_model=64, n_heads=4, conv_kernel_size=15, ff_expansion_factor=2), 8, 2, 90,
                    [90, 41], False),
    "k9_ff8_d176": (dict(feat_in=80, n_layers=1, d_model=176, n_heads=4, conv_kernel_size=9, ff_expansion_factor=8), 9, 2, 120,
                    [120, 77], False),
    "tied_d256": (dict(feat_in=80, n_layers=3, d_model=256, n_heads=4, untie_biases=False), 10, 2, 130, [130, 64], False),
}


def checksums(sd):
    return {k: float(v.double().sum()) for k, v in sd.items()}


def main():
    torch.set_num_threads(1)  # fixed reduction order
    only = set(sys.argv[1:])
    for name, (kw, seed, b, t, lengths, store) in CASES.items():
        if only and name not in only:
            continue
        cfg = EncoderConfig(**kw)
        sd = random_state_dict(cfg, seed)
        enc = build_reference_encoder(cfg, sd)
        x, length = synthetic_batch(b, cfg.feat_in, t, lengths, seed=1234 + seed)
        with torch.no_grad():
            # NOTE: the reference's own length=None branch (conformer_encoder.py:243-246) calls
            # Tensor.new_full(int, ...) which raises TypeError on torch >= 2.x; its documented meaning is
            # "every row is full length", so the no-length case is generated with an explicit full-length vector.
            y, ylen = enc(audio_signal=x, length=length)
        arrays = {
            "audio_signal": x.numpy(),
            "length": length.numpy(),
            "has_length": np.array(lengths is not None),
            "encoded": y.contiguous().numpy(),
            "encoded_len": ylen.numpy(),
            "meta": np.array(json.dumps({"config": dataclasses.asdict(cfg), "weight_seed": seed,
                                         "checksums": checksums(sd), "torch": torch.__version__})),
        }
        if store:
            for k, v in sd.items():
                arrays["w:" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        print(name, tuple(y.shape), ylen.tolist())

    if only:
        return
    # calc_length pins (reference function, float32 arithmetic inside)
    from nemo.collections.asr.parts.submodules.subsampling import calc_length

    vals = torch.tensor([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 333, 640, 900, 1000, 2000, 2001, 29999, 30000,
                         (1 << 24) - 1, 1 << 24, (1 << 24) + 1, (1 << 24) + 3, (1 << 25) + 5], dtype=torch.int64)
    out = {"lengths": vals.numpy()}
    for rep in (1, 2, 3):
        out[f"rep{rep}"] = calc_length(vals, padding=1, kernel_size=3, stride=2, ceil_mode=False,
                                       repeat_num=rep).numpy()
    np.savez_compressed(os.path.join(OUT, "lengths.npz"), **out)
    print("lengths", out["rep2"].tolist())


if __name__ == "__main__":
    main()
