"""Generates tests/golden/rnnt_*.npz by running the UNMODIFIED reference RNNTDecoder / RNNTJoint /
GreedyBatchedRNNTInfer (modules/rnnt.py, parts/submodules/rnnt_greedy_decoding.py, loaded through
oracle/reference_loader.py) on seeded weights and encoder outputs.  Build container only:

    python tests/golden/make_golden_rnnt.py
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference_rnnt_classes  # noqa: E402
from oracle.rnnt_oracle import random_rnnt_state_dicts  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {  # name: (enc_hidden, pred_hidden, joint_hidden, vocab, B, T, lens, max_symbols, activation, blank_bias, seed)
    "rnnt_tiny": (48, 64, 56, 30, 5, 40, [40, 0, 17, 1, 33], 5, "relu", 0.45, 0),
    "rnnt_bpe128": (64, 96, 80, 128, 9, 60, [60, 60, 51, 44, 38, 30, 22, 9, 3], 3, "relu", 0.6, 1),
    "rnnt_tanh": (32, 40, 48, 20, 4, 50, [50, 41, 50, 12], 8, "tanh", 0.5, 2),
    "rnnt_dense_emit": (48, 64, 56, 30, 3, 25, [25, 20, 11], 4, "relu", 0.0, 3),
}


def run_reference(dec_sd, joint_sd, dims, x, lens, max_symbols, activation):
    enc_hidden, pred_hidden, joint_hidden, vocab = dims
    dec_cls, joint_cls, greedy_cls = load_reference_rnnt_classes()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dec = dec_cls(prednet=dict(pred_hidden=pred_hidden, pred_rnn_layers=1, dropout=0.1), vocab_size=vocab)
    joint = joint_cls(jointnet=dict(encoder_hidden=enc_hidden, pred_hidden=pred_hidden, joint_hidden=joint_hidden,
                                    activation=activation, dropout=0.1), num_classes=vocab)
    dec.load_state_dict(dec_sd, strict=True)
    joint.load_state_dict(joint_sd, strict=True)
    greedy = greedy_cls(dec, joint, blank_index=vocab, max_symbols_per_step=max_symbols)
    (hyps,) = greedy(encoder_output=x, encoded_lengths=lens)
    return hyps


def main():
    torch.set_num_threads(1)
    for name, (e, p, j, v, b, t, lens, max_symbols, activation, blank_bias, seed) in CASES.items():
        dec_sd, joint_sd = random_rnnt_state_dicts(e, p, j, v, seed, blank_bias)
        x = torch.randn(b, e, t, generator=torch.Generator().manual_seed(100 + seed))
        lens_t = torch.tensor(lens, dtype=torch.int64)
        hyps = run_reference(dec_sd, joint_sd, (e, p, j, v), x, lens_t, max_symbols, activation)
        n = [len(h.y_sequence) for h in hyps]
        tokens = np.full((b, max(max(n), 1)), -1, dtype=np.int64)
        steps = np.full_like(tokens, -1)
        hs = np.zeros((b, p), dtype=np.float32)
        cs = np.zeros((b, p), dtype=np.float32)
        for i, h in enumerate(hyps):
            tokens[i, :n[i]] = h.y_sequence.numpy()
            steps[i, :n[i]] = np.asarray(h.timestep, dtype=np.int64)
            if h.dec_state is not None:
                hs[i] = h.dec_state[0][0].numpy()
                cs[i] = h.dec_state[1][0].numpy()
        arrays = dict(encoder_output=x.numpy(), encoded_lengths=lens_t.numpy(), dims=np.array([e, p, j, v]),
                      max_symbols=np.array(-1 if max_symbols is None else max_symbols), activation=np.array(activation),
                      blank_bias=np.array(blank_bias), weight_seed=np.array(seed), n_tokens=np.array(n), tokens=tokens,
                      timesteps=steps, scores=np.array([h.score for h in hyps], dtype=np.float64), h=hs, c=cs,
                      has_state=np.array(hyps[0].dec_state is not None),
                      weight_checksum=np.array(float(sum(w.double().sum() for w in list(dec_sd.values()) + list(joint_sd.values())))))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        print(name, "tokens per utterance", n, "frames", lens)


if __name__ == "__main__":
    main()
