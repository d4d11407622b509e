"""Generates tests/golden/ctc_collapse.npz by running the UNMODIFIED reference greedy CTC collapse
(``WER.ctc_decoder_predictions_tensor``, metrics/wer.py:122-188, loaded through oracle/reference_loader.py) on seeded
arg-max frames.  Build container only:

    python tests/golden/make_golden_ctc_collapse.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.reference_loader import reference_ctc_collapse  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    arrays = {}
    n = 0
    for seed, (b, t, v, p_blank, p_repeat) in enumerate([(4, 40, 28, 0.5, 0.3), (3, 97, 128, 0.7, 0.5), (6, 13, 5, 0.2, 0.6),
                                                         (2, 200, 1024, 0.8, 0.2), (5, 1, 3, 0.5, 0.5)]):
        g = torch.Generator().manual_seed(seed)
        pred = torch.randint(0, v, (b, t), generator=g)
        for row in range(b):  # speech-like: many blanks, runs of repeats
            for i in range(t):
                r = float(torch.rand((), generator=g))
                if r < p_blank:
                    pred[row, i] = v
                elif r < p_blank + p_repeat * (1 - p_blank) and i > 0:
                    pred[row, i] = pred[row, i - 1]
        lens = torch.randint(0, t + 1, (b,), generator=g)
        lens[0] = t
        for tag, ln in (("len", lens), ("nolen", None)):
            for fold in (True, False):
                out = reference_ctc_collapse(pred, ln, v, fold)
                width = max(max((len(o) for o in out), default=0), 1)
                tok = np.full((b, width), -1, dtype=np.int64)
                for i, o in enumerate(out):
                    tok[i, :len(o)] = o
                key = f"c{n}"
                arrays[key + "_pred"] = pred.numpy()
                arrays[key + "_lens"] = ln.numpy() if ln is not None else np.zeros(0, dtype=np.int64)
                arrays[key + "_blank"] = np.array(v)
                arrays[key + "_fold"] = np.array(fold)
                arrays[key + "_tokens"] = tok
                arrays[key + "_count"] = np.array([len(o) for o in out])
                n += 1
    arrays["n_cases"] = np.array(n)
    np.savez_compressed(os.path.join(OUT, "ctc_collapse.npz"), **arrays)
    print("ctc_collapse", n, "cases")


if __name__ == "__main__":
    main()
