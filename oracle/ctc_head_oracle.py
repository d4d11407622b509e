"""CPU oracle for the CTC head that consumes the encoder output -- TEST INFRASTRUCTURE ONLY (imported by tests/ and
tests/golden/make_golden_ctc.py; the product path never touches it).

Restates, in plain torch fp32:
  * ConvASRDecoder.forward      nemo/collections/asr/modules/conv_asr.py:437-444
        log_softmax(Conv1d(feat_in, num_classes + 1, kernel_size=1)(encoder_output).transpose(1, 2), dim=-1)
  * the greedy predictions      nemo/collections/asr/models/ctc_models.py:593-594   log_probs.argmax(dim=-1)
  * the greedy CTC collapse     nemo/collections/asr/metrics/wer.py:152-170         fold repeats, drop blanks
Pinned against the reference's own ConvASRDecoder (loaded unmodified by oracle/reference_loader.py) through
tests/golden/ctc_head_*.npz; the collapse loop has no importable reference here (wer.py needs editdistance /
torchmetrics) and is pinned by hand-written cases only.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


def random_head_state_dict(feat_in: int, num_classes: int, seed: int) -> Dict[str, torch.Tensor]:
    """Weights with the reference's parameter names and shapes (xavier_uniform like init_weights, conv_asr.py:435)."""
    g = torch.Generator().manual_seed(seed)
    v1 = num_classes + 1  # + blank (conv_asr.py:431)
    bound = (6.0 / (feat_in + v1)) ** 0.5
    return {
        "decoder_layers.0.weight": (torch.rand(v1, feat_in, 1, generator=g) * 2 - 1) * bound,
        "decoder_layers.0.bias": (torch.rand(v1, generator=g) * 2 - 1) * 0.1,
    }


def ctc_head_forward(sd: Dict[str, torch.Tensor], encoder_output: torch.Tensor) -> torch.Tensor:
    """encoder_output (B, D, T) -> log_probs (B, T, V+1)   (conv_asr.py:437-438)."""
    y = F.conv1d(encoder_output.float(), sd["decoder_layers.0.weight"].float(), sd["decoder_layers.0.bias"].float())
    return F.log_softmax(y.transpose(1, 2), dim=-1)


def greedy_predictions(log_probs: torch.Tensor) -> torch.Tensor:
    """ctc_models.py:594."""
    return log_probs.argmax(dim=-1, keepdim=False)


def greedy_collapse(predictions: torch.Tensor, lengths: Optional[Sequence[int]], blank_id: int) -> List[List[int]]:
    """wer.py:152-164: per utterance, cut at its length, keep p when it differs from its predecessor (or the
    predecessor is blank) and is not blank."""
    out = []
    for b in range(predictions.shape[0]):
        seq = predictions[b].tolist()
        if lengths is not None:
            seq = seq[: int(lengths[b])]
        decoded, previous = [], blank_id
        for p in seq:
            if (p != previous or previous == blank_id) and p != blank_id:
                decoded.append(p)
            previous = p
        out.append(decoded)
    return out
