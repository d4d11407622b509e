"""CTC head on the B200 encoder output: drop-in for ``nemo.collections.asr.modules.ConvASRDecoder``
(modules/conv_asr.py:397-444) plus the greedy step the CTC models take right after it.

Same constructor arguments, same ``state_dict`` keys (``decoder_layers.0.weight`` (V+1, feat_in, 1),
``decoder_layers.0.bias`` (V+1)), same call ``decoder(encoder_output=(B, D, T)) -> log_probs (B, T, V+1)``.
The projection runs on the tcgen05 GEMM of libcfb.so (bf16 operands, fp32 logits), log_softmax and the argmax in
one memory-bound kernel (csrc/ctc_head.cu).  No CPU fallback.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib


class ConvASRDecoder(nn.Module):
    def __init__(self, feat_in: int, num_classes: int, init_mode: str = "xavier_uniform", vocabulary=None):
        super().__init__()
        if vocabulary is None and num_classes < 0:  # conv_asr.py:416-419
            raise ValueError("Neither of the vocabulary and num_classes are set! At least one of them need to be set.")
        if num_classes <= 0:
            num_classes = len(vocabulary)
        if vocabulary is not None:
            if num_classes != len(vocabulary):  # conv_asr.py:425-428
                raise ValueError("If vocabulary is specified, it's length should be equal to the num_classes. "
                                 f"Instead got: num_classes={num_classes} and len(vocabulary)={len(vocabulary)}")
            self._vocabulary = list(vocabulary)
        if feat_in % 8 != 0:
            raise NotImplementedError("ConvASRDecoder (B200): feat_in must be a multiple of 8")
        self._feat_in = feat_in
        self._num_classes = num_classes + 1  # + blank (conv_asr.py:431)
        self.decoder_layers = nn.Sequential(nn.Conv1d(feat_in, self._num_classes, kernel_size=1, bias=True))
        if init_mode == "xavier_uniform":  # parts/submodules/jasper.py:97-100
            nn.init.xavier_uniform_(self.decoder_layers[0].weight, gain=1.0)
        self._packed = None
        self._scratch = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())
        self.eval()

    # ---- reference surface
    @property
    def input_types(self):
        return OrderedDict({"encoder_output": ("B", "D", "T")})

    @property
    def output_types(self):
        return OrderedDict({"logprobs": ("B", "T", "D")})

    @property
    def vocabulary(self):
        return getattr(self, "_vocabulary", None)

    @property
    def num_classes_with_blank(self):
        return self._num_classes

    def input_example(self, max_batch=1, max_dim=256):
        return (torch.randn(max_batch, self._feat_in, max_dim, device=next(self.parameters()).device),)

    def _invalidate(self):
        self._packed = None

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_packed"] = None
        return super()._apply(fn, *args, **kwargs)

    # ---- forward
    def _prepare(self, device):
        w = self.decoder_layers[0].weight.detach()[:, :, 0].to(device=device, dtype=torch.bfloat16).contiguous()
        b = self.decoder_layers[0].bias.detach().to(device=device, dtype=torch.float32).contiguous()
        self._packed = (w, b, device)

    @torch.no_grad()
    def forward_with_predictions(self, encoder_output: torch.Tensor, _int32_predictions: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """(log_probs (B, T, V+1) fp32, greedy predictions (B, T) int64 = log_probs.argmax(-1), ctc_models.py:594)."""
        if encoder_output.dim() != 3 or encoder_output.size(1) != self._feat_in:
            raise TypeError(f"encoder_output must be (B, {self._feat_in}, T), got {tuple(encoder_output.shape)}")
        if not encoder_output.is_cuda:
            raise RuntimeError("ConvASRDecoder (B200) has no CPU path: encoder_output must be a CUDA tensor")
        device = encoder_output.device
        if self._packed is None or self._packed[2] != device:
            self._prepare(device)
        w, bias, _ = self._packed
        b, d, t = encoder_output.shape
        x = encoder_output.transpose(1, 2)  # the encoder returns the transposed view of a contiguous (B, T, D) buffer
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        m, v1 = b * t, self._num_classes
        lib = _lib.load_library()
        need = lib.cfb_ctc_head_scratch_bytes(m, d, v1) + 256
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != device:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=device)
        sptr = (self._scratch.data_ptr() + 255) // 256 * 256
        log_probs = torch.empty(b, t, v1, dtype=torch.float32, device=device)
        best = torch.empty(b, t, dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            rc = lib.cfb_op_ctc_head(ctypes.c_void_p(x.data_ptr()), _lib.CFB_F32 if x.dtype == torch.float32 else _lib.CFB_BF16,
                                     ctypes.c_void_p(w.data_ptr()), ctypes.c_void_p(bias.data_ptr()), m, d, v1,
                                     ctypes.c_void_p(log_probs.data_ptr()), ctypes.c_void_p(best.data_ptr()),
                                     ctypes.c_void_p(sptr), need - 256,
                                     ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        _lib.check(rc, None, "cfb_op_ctc_head")
        return log_probs, (best if _int32_predictions else best.long())

    def forward(self, encoder_output: torch.Tensor) -> torch.Tensor:
        return self.forward_with_predictions(encoder_output)[0]

    @torch.no_grad()
    def greedy_tokens(self, encoder_output: torch.Tensor, encoded_lengths: Optional[torch.Tensor] = None):
        """features -> token ids without leaving the device: head, arg-max and the greedy collapse (metrics/wer.py:152-164).
        Returns (tokens (B, T) int32, n_tokens (B) int32), both on the device; row b holds n_tokens[b] ids."""
        _, pred = self.forward_with_predictions(encoder_output, _int32_predictions=True)
        return ctc_collapse_device(pred, encoded_lengths, self._num_classes - 1)


def ctc_collapse_device(predictions: torch.Tensor, lengths: Optional[torch.Tensor], blank_id: int):
    """Greedy CTC collapse on the device (cfb_op_ctc_collapse): predictions (B, T) CUDA integer tensor, lengths (B) or None
    -> (tokens (B, T) int32, n_tokens (B) int32) on the device.  Enqueue-only."""
    if not predictions.is_cuda or predictions.dim() != 2:
        raise RuntimeError("ctc_collapse_device needs a (B, T) CUDA tensor of arg-max classes")
    device = predictions.device
    pred = predictions.to(torch.int32).contiguous()
    lens = None if lengths is None else lengths.to(device=device, dtype=torch.int32).contiguous()
    b, t = pred.shape
    tokens = torch.empty(b, t, dtype=torch.int32, device=device)
    n_tokens = torch.empty(b, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        rc = _lib.load_library().cfb_op_ctc_collapse(
            ctypes.c_void_p(pred.data_ptr()), ctypes.c_void_p(lens.data_ptr()) if lens is not None else None, b, t, int(blank_id),
            ctypes.c_void_p(tokens.data_ptr()), ctypes.c_void_p(n_tokens.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    _lib.check(rc, None, "cfb_op_ctc_collapse")
    return tokens, n_tokens


def ctc_greedy_decode(predictions: torch.Tensor, lengths: Optional[Sequence[int]], blank_id: int) -> List[List[int]]:
    """The greedy CTC collapse of metrics/wer.py:152-164: per utterance cut at its length, fold consecutive repeats, drop
    blanks.  CUDA predictions are collapsed on the device (one kernel, then one read of the ids); CPU predictions by the
    reference's own host loop."""
    if predictions.is_cuda:
        lens_t = None if lengths is None else (lengths if torch.is_tensor(lengths) else torch.tensor(list(lengths)))
        tokens, n_tokens = ctc_collapse_device(predictions, lens_t, blank_id)
        n = n_tokens.cpu().tolist()
        host = tokens[:, :max(max(n), 1)].cpu()
        return [host[i, :k].tolist() for i, k in enumerate(n)]
    pred = predictions.long().cpu()
    lens = None if lengths is None else [int(v) for v in (lengths.cpu().tolist() if torch.is_tensor(lengths) else lengths)]
    out = []
    for b in range(pred.shape[0]):
        seq = pred[b].tolist()
        if lens is not None:
            seq = seq[: lens[b]]
        decoded, previous = [], blank_id
        for p in seq:
            if (p != previous or previous == blank_id) and p != blank_id:
                decoded.append(p)
            previous = p
        out.append(decoded)
    return out
