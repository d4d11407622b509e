"""Generates tests/golden/ctc_head_*.npz by running the UNMODIFIED reference ConvASRDecoder (modules/conv_asr.py,
loaded through oracle/reference_loader.py) on seeded weights and encoder outputs.  Build container only:

    python tests/golden/make_golden_ctc.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ctc_head_oracle import random_head_state_dict  # noqa: E402
from oracle.reference_loader import load_reference_ctc_decoder_class  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {  # name: (feat_in, num_classes, B, T, seed)
    "ctc_head_char_d176": (176, 28, 2, 37, 0),
    "ctc_head_bpe_d256": (256, 128, 3, 50, 1),
    "ctc_head_bpe1024_d512": (512, 1024, 2, 21, 2),
}


def main():
    torch.set_num_threads(1)
    cls = load_reference_ctc_decoder_class()
    for name, (d, v, b, t, seed) in CASES.items():
        sd = random_head_state_dict(d, v, seed)
        dec = cls(feat_in=d, num_classes=v)
        dec.load_state_dict(sd, strict=True)
        dec.eval()
        x = torch.randn(b, d, t, generator=torch.Generator().manual_seed(100 + seed))
        with torch.no_grad():
            lp = dec(encoder_output=x)
        arrays = dict(encoder_output=x.numpy(), log_probs=lp.numpy(), num_classes=np.array(v), weight_seed=np.array(seed),
                      weight_checksum=np.array(float(sd["decoder_layers.0.weight"].double().sum())))
        if d * v <= 8192:  # small heads carry their weights; the others are re-created from the seed and checked
            arrays.update(weight=sd["decoder_layers.0.weight"].numpy(), bias=sd["decoder_layers.0.bias"].numpy())
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        print(name, tuple(lp.shape), float(lp.exp().sum(-1).mean()))


if __name__ == "__main__":
    main()
