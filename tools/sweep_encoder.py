"""Randomised parity sweep of the CUDA encoder (run on the GPU box): random recipes (d_model, heads, depth, depth-wise kernel,
expansion factor, subsampling channels, tied / untied biases, xscaling, feat_out), batch sizes, extents and lengths (1 frame
to full length, one zero-length row now and then), against

  * the CPU oracle (tolerances of BASELINE.json: encoded_len bit-exact; rel-L2 <= 1e-2, max-abs <= 5e-2 on valid frames, bf16),
  * itself: the packed forward must equal the dense forward BIT FOR BIT on every valid frame, frames behind encoded_len are
    exact zeros, and a second run of the same call repeats the first bit for bit.

    python tools/sweep_encoder.py [n_cases] [seed]

Prints one line per failing case and a JSON summary; exit code 1 on any failure."""
import json
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import conformer_nemo_b200 as cn  # noqa: E402
from oracle import conformer_oracle as oc  # noqa: E402  (the checker)


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad, worst_l2, worst_abs, frames, packed_cases = [], 0.0, 0.0, 0, 0
    for case in range(n_cases):
        d, h = rnd.choice([(64, 4), (64, 2), (176, 4), (256, 4), (256, 8), (512, 8), (512, 16), (144, 4)])
        kw = dict(feat_in=80, n_layers=rnd.choice([1, 2, 2, 3]), d_model=d, n_heads=h,
                  conv_kernel_size=rnd.choice([31, 31, 31, 15, 9, 5]), ff_expansion_factor=rnd.choice([4, 4, 4, 2, 8]),
                  subsampling_conv_channels=rnd.choice([-1, -1, 64, 128]), xscaling=rnd.random() < 0.8,
                  untie_biases=rnd.random() < 0.7, feat_out=rnd.choice([-1, -1, -1, 96]))
        cfg = oc.EncoderConfig(**kw)
        sd = oc.random_state_dict(cfg, 500 + case)
        b = rnd.choice([1, 2, 3, 5, 8, 9, 12])
        t = rnd.randint(9, 900)
        lens = [rnd.randint(1, t) for _ in range(b)]
        lens[rnd.randrange(b)] = t
        if b > 2 and rnd.random() < 0.2:
            lens[rnd.randrange(b)] = 0 if lens.count(t) > 1 or lens[0] != t else lens[-1]
        if t not in lens:
            lens[0] = t
        x, length = oc.synthetic_batch(b, 80, t, lens, seed=900 + case)
        want, want_len = oc.encoder_forward(sd, cfg, x, length)
        enc = cn.ConformerEncoder(feat_in=80, n_layers=cfg.n_layers, d_model=d, n_heads=h, feat_out=cfg.feat_out,
                                  subsampling_conv_channels=cfg.subsampling_conv_channels,
                                  ff_expansion_factor=cfg.ff_expansion_factor, xscaling=cfg.xscaling,
                                  conv_kernel_size=cfg.conv_kernel_size, untie_biases=cfg.untie_biases)
        enc.load_state_dict(sd, strict=False)
        enc = enc.cuda().eval()
        enc.packed = False
        y, ylen = enc(audio_signal=x.cuda(), length=length.cuda())
        y, ylen = y.clone(), ylen.clone()
        y2, _ = enc(audio_signal=x.cuda(), length=length.cuda())
        why = []
        if not torch.equal(y, y2):
            why.append("not repeatable")
        if not torch.equal(ylen.cpu(), want_len):
            why.append(f"encoded_len {ylen.tolist()} != {want_len.tolist()}")
        yc = y.float().cpu()
        for row, n in enumerate(want_len.tolist()):
            if n < yc.shape[2] and float(yc[row, :, n:].abs().max()) != 0.0:
                why.append(f"row {row}: non-zero behind encoded_len")
            if n > 0:
                g, w = yc[row, :, :n].double(), want[row, :, :n].double()
                l2, mx = float((g - w).norm() / w.norm()), float((g - w).abs().max())
                worst_l2, worst_abs, frames = max(worst_l2, l2), max(worst_abs, mx), frames + n
                if not (l2 <= 1e-2 and mx <= 5e-2):
                    why.append(f"row {row}: rel_l2 {l2:.3e} max_abs {mx:.3e}")
        if cfg.feat_out == -1:  # the packed forward serves encoders without out_proj
            enc.packed = True
            yp, ylp = enc(audio_signal=x.cuda(), length=length.cuda(), length_host=lens)
            packed_cases += 1
            if not torch.equal(ylp, ylen):
                why.append("packed: encoded_len differs")
            for row, n in enumerate(want_len.tolist()):
                if not torch.equal(yp[row, :, :n], y[row, :, :n]):
                    why.append(f"packed: row {row} differs from dense by {float((yp[row, :, :n].float() - y[row, :, :n].float()).abs().max()):.3e}")
                if n < yp.shape[2] and float(yp[row, :, n:].abs().max()) != 0.0:
                    why.append(f"packed: row {row} non-zero behind encoded_len")
        if why:
            bad.append(dict(case=case, config=kw, B=b, T=t, lens=lens, why=why))
            print("FAIL", json.dumps(bad[-1]))
        del enc
    print(json.dumps(dict(cases=n_cases, packed_cases=packed_cases, valid_frames=frames, failures=len(bad),
                          worst_rel_l2=worst_l2, worst_max_abs=worst_abs)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
