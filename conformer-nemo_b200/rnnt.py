"""Transducer decoder, joint and greedy-batch inference on the B200 encoder output: drop-ins for
``nemo.collections.asr.modules.RNNTDecoder`` (modules/rnnt.py:51-611), ``nemo.collections.asr.modules.RNNTJoint``
(modules/rnnt.py:613-1084) and ``GreedyBatchedRNNTInfer`` (parts/submodules/rnnt_greedy_decoding.py:358-616) for
inference with ``decoding.strategy = greedy_batch`` (configs/conformer_transducer_bpe.yaml:139-177).

Same constructor arguments, same ``state_dict`` keys (``prediction.embed.weight``, ``prediction.dec_rnn.lstm.*_l0``,
``pred.*``, ``enc.*``, ``joint_net.{n}.*``), same call
``greedy(encoder_output=(B, D, T), encoded_lengths=(B,)) -> (List[Hypothesis],)`` with ``y_sequence`` / ``timestep`` /
``score`` / ``dec_state`` / ``length`` filled like the reference does.  The whole decode -- joint.enc over every frame,
then the symbol loop with the LSTM and the joint -- runs in libcfb.so (``cfb_op_rnnt_greedy``, csrc/rnnt_greedy.cu)
with ONE device-to-host read at the end; the reference synchronises once per symbol step.  No CPU fallback; training
(``forward`` over label sequences, the transducer loss) stays with the reference modules.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from . import _lib

_ACTIVATIONS = {"relu": 0, "sigmoid": 1, "tanh": 2}


@dataclass
class Hypothesis:
    """The fields of parts/utils/rnnt_utils.py:33-82 that greedy decoding fills."""
    score: float
    y_sequence: Union[List[int], torch.Tensor]
    text: Optional[str] = None
    dec_out: Optional[List[torch.Tensor]] = None
    dec_state: Optional[Any] = None
    timestep: Union[List[int], torch.Tensor] = field(default_factory=list)
    alignments: Optional[Any] = None
    length: Union[int, torch.Tensor] = 0
    y: Optional[List[torch.Tensor]] = None
    lm_state: Optional[Any] = None
    lm_scores: Optional[torch.Tensor] = None
    tokens: Optional[Any] = None
    last_token: Optional[torch.Tensor] = None


class _LSTMDropout(nn.Module):
    """Parameter container with the key layout of common/parts/rnn.py:151-226 (``lstm.weight_ih_l0`` ...)."""

    def __init__(self, input_size, hidden_size, forget_gate_bias, weights_init_scale, hidden_hidden_bias_scale):
        super().__init__()
        self.lstm = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=1)
        if forget_gate_bias is not None:  # rnn.py:212-219
            with torch.no_grad():
                self.lstm.bias_ih_l0[hidden_size:2 * hidden_size].fill_(forget_gate_bias)
                self.lstm.bias_hh_l0[hidden_size:2 * hidden_size] *= float(hidden_hidden_bias_scale)
        with torch.no_grad():
            for p in self.parameters():
                p *= float(weights_init_scale)


class RNNTDecoder(nn.Module):
    def __init__(self, prednet: Dict[str, Any], vocab_size: int, normalization_mode: Optional[str] = None,
                 random_state_sampling: bool = False, blank_as_pad: bool = True):
        super().__init__()
        self.pred_hidden = prednet["pred_hidden"]
        self.pred_rnn_layers = prednet["pred_rnn_layers"]
        self.blank_idx = vocab_size
        self.vocab_size = vocab_size
        self.blank_as_pad = blank_as_pad
        self.random_state_sampling = random_state_sampling
        if normalization_mode not in (None, "batch", "layer"):
            raise ValueError(f"unknown norm={normalization_mode}")  # common/parts/rnn.py:80-81
        unsupported = []
        if self.pred_rnn_layers != 1:
            unsupported.append(f"pred_rnn_layers={self.pred_rnn_layers} (the Conformer-Transducer recipes use 1)")
        if normalization_mode is not None:
            unsupported.append(f"normalization_mode={normalization_mode!r}")
        if not blank_as_pad:
            unsupported.append("blank_as_pad=False (greedy_batch needs the padding-row embedding)")
        if prednet.get("t_max") is not None:
            unsupported.append("t_max (chrono initialisation is a training-time choice)")
        if prednet.get("rnn_hidden_size", -1) not in (-1, self.pred_hidden):
            unsupported.append("rnn_hidden_size != pred_hidden (projected LSTM)")
        if self.pred_hidden % 4:
            unsupported.append("pred_hidden not a multiple of 4")
        if unsupported:
            raise NotImplementedError("RNNTDecoder (B200): " + "; ".join(unsupported))
        self.prediction = nn.ModuleDict({
            "embed": nn.Embedding(vocab_size + 1, self.pred_hidden, padding_idx=self.blank_idx),  # rnnt.py:327
            "dec_rnn": _LSTMDropout(self.pred_hidden, self.pred_hidden, prednet.get("forget_gate_bias", 1.0),
                                    prednet.get("weights_init_scale", 1.0), prednet.get("hidden_hidden_bias_scale", 0.0)),
        })
        self.eval()

    def forward(self, *args, **kwargs):
        raise NotImplementedError("RNNTDecoder (B200) serves greedy inference through GreedyBatchedRNNTInfer; the "
                                  "teacher-forced forward over label sequences is training code and stays with the reference")


class RNNTJoint(nn.Module):
    def __init__(self, jointnet: Dict[str, Any], num_classes: int, vocabulary: Optional[List] = None,
                 log_softmax: Optional[bool] = None, preserve_memory: bool = False, fuse_loss_wer: bool = False,
                 fused_batch_size: Optional[int] = None, experimental_fuse_loss_wer: Any = None):
        super().__init__()
        self.vocabulary = vocabulary
        self._vocab_size = num_classes
        self._num_classes = num_classes + 1  # + blank (rnnt.py:741)
        if experimental_fuse_loss_wer is not None:
            fuse_loss_wer = experimental_fuse_loss_wer
        if fuse_loss_wer and fused_batch_size is None:  # rnnt.py:754-755
            raise ValueError("If `fuse_loss_wer` is set, then `fused_batch_size` cannot be None!")
        self.log_softmax = log_softmax
        self.encoder_hidden = jointnet["encoder_hidden"]
        self.pred_hidden = jointnet["pred_hidden"]
        self.joint_hidden = jointnet["joint_hidden"]
        self.activation = str(jointnet["activation"]).lower()
        if self.activation not in _ACTIVATIONS:  # rnnt.py:1023-1024
            raise ValueError("Unsupported activation for joint step - please pass one of [relu, sigmoid, tanh]")
        if self.encoder_hidden % 8 or self.joint_hidden % 4:
            raise NotImplementedError("RNNTJoint (B200): encoder_hidden must be a multiple of 8 and joint_hidden of 4")
        dropout = jointnet.get("dropout", 0.0)
        self.pred = nn.Linear(self.pred_hidden, self.joint_hidden)
        self.enc = nn.Linear(self.encoder_hidden, self.joint_hidden)
        act = {"relu": nn.ReLU(inplace=True), "sigmoid": nn.Sigmoid(), "tanh": nn.Tanh()}[self.activation]
        layers = [act] + ([nn.Dropout(p=dropout)] if dropout else []) + [nn.Linear(self.joint_hidden, self._num_classes)]
        self.joint_net = nn.Sequential(*layers)  # the output Linear sits at index 2 when dropout != 0, like rnnt.py:1039-1043
        self.eval()

    @property
    def num_classes_with_blank(self):
        return self._num_classes

    def forward(self, *args, **kwargs):
        raise NotImplementedError("RNNTJoint (B200) serves greedy inference through GreedyBatchedRNNTInfer; the (B, T, U, V) "
                                  "joint tensor and the fused loss / WER are training code and stay with the reference")


class GreedyBatchedRNNTInfer:
    """rnnt_greedy_decoding.py:358-616 (the ``blank_as_pad`` branch)."""

    def __init__(self, decoder_model: RNNTDecoder, joint_model: RNNTJoint, blank_index: int,
                 max_symbols_per_step: Optional[int] = None, preserve_alignments: bool = False):
        if preserve_alignments:
            raise NotImplementedError("GreedyBatchedRNNTInfer (B200): preserve_alignments is a debugging aid of the reference")
        if blank_index != decoder_model.blank_idx or blank_index != joint_model.num_classes_with_blank - 1:
            raise ValueError("blank_index must be len(vocabulary) for the decoder and the joint (blank_as_pad)")
        if decoder_model.pred_hidden != joint_model.pred_hidden:
            raise ValueError("decoder pred_hidden and joint pred_hidden differ")
        self.decoder = decoder_model
        self.joint = joint_model
        self._blank_index = blank_index
        self._SOS = blank_index
        self.max_symbols = max_symbols_per_step
        self.preserve_alignments = False
        self._packed = None
        self._scratch = None

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def invalidate(self):
        """Call after changing decoder / joint parameters (they are re-read on the next call)."""
        self._packed = None

    def _sources(self):
        dec, jn = self.decoder, self.joint
        lstm = dec.prediction["dec_rnn"].lstm
        out = [m for m in jn.joint_net if isinstance(m, nn.Linear)][-1]
        return (dec.prediction["embed"].weight, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0,
                jn.pred.weight, jn.pred.bias, jn.enc.weight, jn.enc.bias, out.weight, out.bias)

    def _fingerprint(self):
        """Identity of the source parameters: storage, dtype and in-place version counter.  The kernel decodes from fp32
        COPIES of them, so ``decoder.half()``, ``.to(dtype)``, re-assigning a Parameter or an in-place update (optimizer
        step, ``load_state_dict``) must refresh the copies -- checked on every call (eleven tuple compares)."""
        return tuple((p.data_ptr(), p.dtype, p._version) for p in self._sources())

    def _prepare(self, device):
        f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        dec, jn = self.decoder, self.joint
        lstm = dec.prediction["dec_rnn"].lstm
        out = [m for m in jn.joint_net if isinstance(m, nn.Linear)][-1]
        t = dict(embed=f(dec.prediction["embed"].weight), w_ih=f(lstm.weight_ih_l0), w_hh=f(lstm.weight_hh_l0),
                 b_ih=f(lstm.bias_ih_l0), b_hh=f(lstm.bias_hh_l0), w_pred=f(jn.pred.weight), b_pred=f(jn.pred.bias),
                 w_enc=f(jn.enc.weight), b_enc=f(jn.enc.bias), w_out=f(out.weight), b_out=f(out.bias))
        if float(t["embed"][self._blank_index].abs().max()) != 0.0:
            raise NotImplementedError("GreedyBatchedRNNTInfer (B200): the blank row of prediction.embed must be the zero padding row")
        w = _lib.CfbRnntWeights()
        w.enc_hidden, w.pred_hidden, w.joint_hidden = jn.encoder_hidden, dec.pred_hidden, jn.joint_hidden
        w.num_classes_with_blank, w.activation = jn.num_classes_with_blank, _ACTIVATIONS[jn.activation]
        for k, v in t.items():
            setattr(w, k, ctypes.cast(ctypes.c_void_p(v.data_ptr()), ctypes.POINTER(ctypes.c_float)))
        self._packed = (w, t, device, self._fingerprint())

    @torch.no_grad()
    def decode_arrays(self, encoder_output: torch.Tensor, encoded_lengths: torch.Tensor, max_tokens: Optional[int] = None):
        """Device-side result: dict of tokens / timesteps (B, max_tokens) int32, n_tokens (B) int32, scores (B) fp32,
        h / c (B, H) fp32.  Enqueue-only (no synchronisation)."""
        jn = self.joint
        if encoder_output.dim() != 3 or encoder_output.size(1) != jn.encoder_hidden:
            raise TypeError(f"encoder_output must be (B, {jn.encoder_hidden}, T), got {tuple(encoder_output.shape)}")
        if not encoder_output.is_cuda:
            raise RuntimeError("GreedyBatchedRNNTInfer (B200) has no CPU path: encoder_output must be a CUDA tensor")
        device = encoder_output.device
        if self._packed is None or self._packed[2] != device or self._packed[3] != self._fingerprint():
            self._prepare(device)
        w = self._packed[0]
        b, d, t = encoder_output.shape
        x = encoder_output.transpose(1, 2)  # the encoder returns the transposed view of a contiguous (B, T, D) buffer
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        lens = encoded_lengths.to(device=device, dtype=torch.int32).contiguous()
        max_symbols = 0 if self.max_symbols is None else int(self.max_symbols)
        if max_tokens is None:
            max_tokens = t * max_symbols if 0 < max_symbols <= 8 else 4 * t + 64
        max_tokens = max(int(max_tokens), 1)
        lib = _lib.load_library()
        need = lib.cfb_rnnt_greedy_scratch_bytes(d, w.pred_hidden, w.joint_hidden, b, t) + 256
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != device:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=device)
        sptr = (self._scratch.data_ptr() + 255) // 256 * 256
        out = dict(tokens=torch.empty(b, max_tokens, dtype=torch.int32, device=device),
                   timesteps=torch.empty(b, max_tokens, dtype=torch.int32, device=device),
                   n_tokens=torch.empty(b, dtype=torch.int32, device=device),
                   scores=torch.empty(b, dtype=torch.float32, device=device),
                   h=torch.empty(b, w.pred_hidden, dtype=torch.float32, device=device),
                   c=torch.empty(b, w.pred_hidden, dtype=torch.float32, device=device),
                   flags=torch.zeros(1, dtype=torch.int32, device=device))
        vp = lambda tensor: ctypes.c_void_p(tensor.data_ptr())
        with torch.cuda.device(device):
            rc = lib.cfb_op_rnnt_greedy(ctypes.byref(w), vp(x), _lib.CFB_F32 if x.dtype == torch.float32 else _lib.CFB_BF16,
                                        vp(lens), b, t, max_symbols, max_tokens, vp(out["tokens"]), vp(out["timesteps"]),
                                        vp(out["n_tokens"]), vp(out["scores"]), vp(out["h"]), vp(out["c"]), vp(out["flags"]),
                                        ctypes.c_void_p(sptr), need - 256,
                                        ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        _lib.check(rc, None, "cfb_op_rnnt_greedy")
        out["_keepalive"] = (x, lens)
        off = sptr - self._scratch.data_ptr()
        out["phase_cycles"] = self._scratch[off + 64:off + 192].view(torch.int64)  # CTA 0's clocks per phase (debugging)
        return out

    @torch.no_grad()
    def forward(self, encoder_output: torch.Tensor, encoded_lengths: torch.Tensor,
                partial_hypotheses: Optional[List[Hypothesis]] = None) -> Tuple[List[Hypothesis]]:
        if partial_hypotheses is not None:  # rnnt_greedy_decoding.py:461-462
            raise NotImplementedError("`partial_hypotheses` support is not supported")
        t = encoder_output.size(2)
        max_tokens = None
        while True:
            out = self.decode_arrays(encoder_output, encoded_lengths, max_tokens)
            n = out["n_tokens"].cpu()  # the one synchronisation of the decode
            flags = int(out["flags"].cpu())
            if flags & 2:
                raise RuntimeError("GreedyBatchedRNNTInfer (B200): runaway emission with max_symbols_per_step=None "
                                   "(4096 symbols at one frame)")
            if flags == 0:
                break
            max_tokens = int(n.max()) + t  # some utterance outgrew the token buffer: run again with room for all of it
        width = max(int(n.max()), 1)
        tokens = out["tokens"][:, :width].cpu().long()
        steps = out["timesteps"][:, :width].cpu()
        scores, h, c = out["scores"].cpu(), out["h"].cpu(), out["c"].cpu()
        lens_cpu = encoded_lengths.to("cpu")
        any_emitted = bool((n > 0).any())
        hyps = []
        for i in range(tokens.size(0)):
            k = int(n[i])
            # pack_hypotheses (rnnt_greedy_decoding.py:43-57): y_sequence as a LongTensor, the state on the CPU; the state is
            # None only when no utterance of the batch ever emitted (`hidden` stays None, :611-614)
            state = ([h[i]], [c[i]]) if any_emitted else None
            hyps.append(Hypothesis(score=float(scores[i]), y_sequence=tokens[i, :k].clone(), timestep=steps[i, :k].tolist(),
                                   dec_state=state, length=lens_cpu[i]))
        return (hyps,)


class GreedyRNNTInfer(GreedyBatchedRNNTInfer):
    """``decoding.strategy = greedy`` (rnnt_greedy_decoding.py:191-355): the reference decodes one utterance at a time with
    the same per-utterance rule, so tokens, timesteps and scores are those of the batched decode (module docstring of
    oracle/rnnt_oracle.py); here the whole batch still runs in one launch.  Differences in what the Hypothesis carries, kept:
    ``dec_state`` is None for an utterance that emitted nothing (:277, :330-346) and ``last_token`` holds the last symbol."""

    @torch.no_grad()
    def forward(self, encoder_output: torch.Tensor, encoded_lengths: torch.Tensor,
                partial_hypotheses: Optional[List[Hypothesis]] = None) -> Tuple[List[Hypothesis]]:
        (hyps,) = super().forward(encoder_output, encoded_lengths, partial_hypotheses)
        for h in hyps:
            if len(h.y_sequence) == 0:
                h.dec_state = None
            else:
                h.last_token = int(h.y_sequence[-1])
        return (hyps,)
