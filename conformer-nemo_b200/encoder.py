"""``ConformerEncoder``: the reference-facing operator.

Same constructor keywords, ``forward(audio_signal, length) -> (encoded, encoded_len)`` contract, attributes and
``state_dict`` key layout as ``nemo.collections.asr.modules.ConformerEncoder``
(nemo/collections/asr/modules/conformer_encoder.py:33-305), so a Hydra config selects it by changing only
``model.encoder._target_`` and existing ``.nemo`` checkpoints load unchanged.  All arithmetic happens in
``libcfb.so`` (hand-written sm_100a CUDA behind the C ABI of include/cfb.h); torch is used for parameter storage,
device buffers and the current stream only.  There is no CPU or eager fallback: ``forward`` raises if the input is
not on a CUDA device or the library is missing.
"""
from __future__ import annotations

import ctypes
import math
from collections import OrderedDict
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib

_DTYPES = {torch.float32: _lib.CFB_F32, torch.bfloat16: _lib.CFB_BF16, torch.float16: _lib.CFB_F16}


class _Holder(nn.Module):
    """Parameter container mirroring a reference sub-module's attribute names; never called."""


def _feed_forward(d_model: int, d_ff: int) -> nn.Module:  # conformer_modules.py:188-193
    m = _Holder()
    m.linear1 = nn.Linear(d_model, d_ff)
    m.linear2 = nn.Linear(d_ff, d_model)
    return m


def _self_attn(d_model: int, n_heads: int, pos_bias_u, pos_bias_v) -> nn.Module:  # multi_head_attention.py:55-157
    m = _Holder()
    for name in ("linear_q", "linear_k", "linear_v", "linear_out"):
        setattr(m, name, nn.Linear(d_model, d_model))
    m.linear_pos = nn.Linear(d_model, d_model, bias=False)
    if pos_bias_u is None or pos_bias_v is None:
        m.pos_bias_u = nn.Parameter(torch.zeros(n_heads, d_model // n_heads))
        m.pos_bias_v = nn.Parameter(torch.zeros(n_heads, d_model // n_heads))
    else:  # untie_biases=False: one pair shared by every layer (conformer_encoder.py:165-173)
        m.pos_bias_u = pos_bias_u
        m.pos_bias_v = pos_bias_v
    return m


def _conv_module(d_model: int, kernel_size: int) -> nn.Module:  # conformer_modules.py:131-158
    m = _Holder()
    m.pointwise_conv1 = nn.Conv1d(d_model, 2 * d_model, 1)
    m.depthwise_conv = nn.Conv1d(d_model, d_model, kernel_size, padding=(kernel_size - 1) // 2, groups=d_model)
    m.batch_norm = nn.BatchNorm1d(d_model)
    m.pointwise_conv2 = nn.Conv1d(d_model, d_model, 1)
    return m


def _layer(d_model, d_ff, n_heads, kernel_size, pos_bias_u, pos_bias_v) -> nn.Module:  # conformer_modules.py:40-86
    m = _Holder()
    m.norm_feed_forward1 = nn.LayerNorm(d_model)
    m.feed_forward1 = _feed_forward(d_model, d_ff)
    m.norm_conv = nn.LayerNorm(d_model)
    m.conv = _conv_module(d_model, kernel_size)
    m.norm_self_att = nn.LayerNorm(d_model)
    m.self_attn = _self_attn(d_model, n_heads, pos_bias_u, pos_bias_v)
    m.norm_feed_forward2 = nn.LayerNorm(d_model)
    m.feed_forward2 = _feed_forward(d_model, d_ff)
    m.norm_out = nn.LayerNorm(d_model)
    return m


def _pre_encode(feat_in: int, d_model: int, channels: int, n_stages: int) -> nn.Module:  # subsampling.py:99-116,151-161
    m = _Holder()
    layers, cin, f = [], 1, feat_in
    for _ in range(n_stages):
        layers += [nn.Conv2d(cin, channels, 3, stride=2, padding=1), nn.ReLU()]
        cin = channels
        f = (f + 2 - 3) // 2 + 1
    m.conv = nn.Sequential(*layers)
    m.out = nn.Linear(channels * f, d_model)
    return m


class ConformerEncoder(nn.Module):
    """B200-native drop-in for the reference ``ConformerEncoder`` (see module docstring).

    Extra keyword (not in the reference): ``precision`` = "bf16" (tcgen05 product path, default) or "fp32_validate"
    (the same dataflow on fp32 CUDA-core kernels, for numerics validation).
    """

    def __init__(
        self,
        feat_in,
        n_layers,
        d_model,
        feat_out=-1,
        subsampling="striding",
        subsampling_factor=4,
        subsampling_conv_channels=-1,
        ff_expansion_factor=4,
        self_attention_model="rel_pos",
        n_heads=4,
        att_context_size=None,
        xscaling=True,
        untie_biases=True,
        pos_emb_max_len=5000,
        conv_kernel_size=31,
        conv_norm_type="batch_norm",
        dropout=0.1,
        dropout_emb=0.1,
        dropout_att=0.0,
        precision="bf16",
    ):
        super().__init__()
        # --- configuration surface: same errors as the reference where it raises, NotImplementedError where the
        #     reference supports something this build does not -- never a silent fallback
        if self_attention_model not in ("rel_pos", "abs_pos"):
            raise ValueError(f"Not valid self_attention_model: '{self_attention_model}'!")  # conformer_encoder.py:191
        if subsampling not in ("striding", "vggnet", "resnet", "subencoder"):
            raise ValueError(f"Not valid sub-sampling: {subsampling}!")  # subsampling.py:149
        if subsampling_factor % 2 != 0:
            raise ValueError("Sampling factor should be a multiply of 2!")  # subsampling.py:63-64
        if conv_norm_type not in ("batch_norm", "layer_norm"):
            raise ValueError(f"conv_norm_type={conv_norm_type} is not valid!")  # conformer_modules.py:153
        if self_attention_model != "rel_pos":
            raise NotImplementedError("only self_attention_model='rel_pos' is built (no shipped config uses abs_pos)")
        if subsampling != "striding" or subsampling_factor != 4:
            raise NotImplementedError("only subsampling='striding' with subsampling_factor=4 is built")
        if conv_norm_type != "batch_norm":
            raise NotImplementedError("only conv_norm_type='batch_norm' is built")
        if att_context_size and tuple(att_context_size) != (-1, -1):
            raise NotImplementedError("limited attention context (att_context_size) is not built")
        if precision not in ("bf16", "fp32_validate"):
            raise ValueError("precision must be 'bf16' or 'fp32_validate'")

        self.d_model = d_model
        self._feat_in = feat_in
        self.n_layers = n_layers
        self.n_heads = n_heads
        self.scale = math.sqrt(d_model)
        self.att_context_size = [-1, -1]
        self.xscale = math.sqrt(d_model) if xscaling else None
        self.pos_emb_max_len = pos_emb_max_len
        self.max_audio_length = pos_emb_max_len  # kept for attribute parity; tables are built per call, sized by T'
        self.use_pad_mask = True
        self.precision = precision
        self._cfg = dict(feat_in=feat_in, n_layers=n_layers, d_model=d_model, feat_out=feat_out,
                         subsampling_factor=subsampling_factor, subsampling_conv_channels=subsampling_conv_channels,
                         ff_expansion_factor=ff_expansion_factor, n_heads=n_heads, conv_kernel_size=conv_kernel_size,
                         xscaling=bool(xscaling))

        channels = d_model if subsampling_conv_channels == -1 else subsampling_conv_channels
        self.pre_encode = _pre_encode(feat_in, d_model, channels, int(math.log(subsampling_factor, 2)))
        if not untie_biases:
            pos_bias_u = nn.Parameter(torch.zeros(n_heads, d_model // n_heads))
            pos_bias_v = nn.Parameter(torch.zeros(n_heads, d_model // n_heads))
        else:
            pos_bias_u = pos_bias_v = None
        self.layers = nn.ModuleList(
            _layer(d_model, d_model * ff_expansion_factor, n_heads, conv_kernel_size, pos_bias_u, pos_bias_v)
            for _ in range(n_layers))
        if feat_out > 0 and feat_out != d_model:
            self.out_proj = nn.Linear(d_model, feat_out)
            self._feat_out = feat_out
        else:
            self.out_proj = None
            self._feat_out = d_model

        self._handle = None        # cfb_handle* (created lazily on the device the parameters live on)
        self._handle_device = None
        self._dirty = True         # parameters changed since the last prepare()
        self._workspace = None
        self._use_graphs = False   # enable_cuda_graphs(): replay one captured graph per input shape
        self._graphs = OrderedDict()
        self._profiling = False
        # "auto": mixed-length batches whose lengths are known on the host run in the packed layout (cfb_forward_packed)
        # when that saves >= 15 % of the token rows; True / False force it on (whenever host lengths exist) / off
        self.packed = "auto"
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.mark_weights_dirty())
        self.eval()

    # ------------------------------------------------------------------------------------------------ reference API
    @property
    def input_types(self):  # conformer_encoder.py:89-98 (neural types are NeMo objects; names and axes are kept)
        return OrderedDict({"audio_signal": ("B", "D", "T"), "length": ("B",)})

    @property
    def output_types(self):  # conformer_encoder.py:100-109
        return OrderedDict({"outputs": ("B", "D", "T"), "encoded_lengths": ("B",)})

    def input_example(self, max_batch=1, max_dim=256):  # conformer_encoder.py:78-87
        dev = next(self.parameters()).device
        return (torch.randn(max_batch, self._feat_in, max_dim, device=dev),
                torch.randint(1, max_dim, (max_batch,), device=dev))

    def freeze(self) -> None:  # nemo/core/classes/module.py:49-56
        for p in self.parameters():
            p.requires_grad = False
        self.eval()

    def unfreeze(self) -> None:  # nemo/core/classes/module.py:58-65
        for p in self.parameters():
            p.requires_grad = True
        self.train()

    def set_max_audio_length(self, max_audio_length):  # conformer_encoder.py:218-229 (no table to grow here)
        self.max_audio_length = max_audio_length

    def update_max_seq_length(self, seq_length: int, device=None):  # conformer_encoder.py:283-294
        # The reference all-reduces the max length so every rank grows its table together; the positional table
        # here is generated per call from T', so no collective and no host sync is needed.
        if seq_length > self.max_audio_length:
            self.max_audio_length = seq_length

    def enable_pad_mask(self, on=True):  # conformer_encoder.py:301-305
        if not on:
            raise NotImplementedError("the CUDA path always applies the padding mask")
        prev, self.use_pad_mask = self.use_pad_mask, on
        return prev

    # ------------------------------------------------------------------------------------------------ CUDA graphs
    def enable_cuda_graphs(self, on: bool = True, max_shapes: int = 4, private_workspaces: bool = False) -> None:
        """cfb_forward only enqueues kernels (no allocation, no synchronisation), so a forward can be captured once
        per input shape and replayed: the ~260 launches of a step then cost one driver call.  With graphs on,
        ``forward`` copies its inputs into the graph's static buffers, replays, and returns views of the graph's
        static output buffers -- they are overwritten by the next forward of the same shape.
        ``private_workspaces``: every captured shape gets its own scratch workspace (instead of sharing the largest),
        which is what lets graphs of different shapes run concurrently on different streams (``forward_many``)."""
        self._use_graphs = bool(on)
        self._graph_cap = max(1, int(max_shapes))
        if bool(private_workspaces) != getattr(self, "_graph_private_ws", False):
            self._graphs.clear()
        self._graph_private_ws = bool(private_workspaces)
        if not on:
            self._graphs.clear()

    @torch.no_grad()
    def forward_many(self, batches, n_streams: int = 3):
        """Independent sub-batches ``[(audio_signal, length), ...]`` (e.g. the length buckets of one rank,
        sharding.plan_shards) run concurrently on up to ``n_streams`` side streams: a small sub-batch leaves most SMs
        idle and its ~250 dependent launches are latency-bound, so overlapping sub-batches hides both.  Needs
        ``enable_cuda_graphs(True, private_workspaces=True)``; otherwise (or for a single sub-batch) runs them one after
        the other.  Returns ``[(encoded, encoded_len), ...]``; the caller's current stream waits for all of them."""
        batches = [tuple(bt) + (None,) * (3 - len(bt)) for bt in batches]  # (audio_signal, length[, length_host])
        if len(batches) <= 1 or n_streams <= 1 or self._profiling or \
                not (self._use_graphs and getattr(self, "_graph_private_ws", False)):
            return [self.forward(audio_signal=x, length=ln, length_host=hl) for x, ln, hl in batches]
        device = batches[0][0].device
        cur = torch.cuda.current_stream(device)
        pool = self.__dict__.setdefault("_side_streams", [])
        while len(pool) < n_streams:
            pool.append(torch.cuda.Stream(device=device))
        fork = torch.cuda.Event()
        fork.record(cur)
        outs, owner = [], {}
        for i, (x, ln, hl) in enumerate(batches):
            # the identity _forward_graph gives the captured forward (same dtype normalisation, same packed decision):
            # two sub-batches with equal keys share one graph and its static buffers, so they must stay on one stream
            in_dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
            host = hl if hl is not None else (ln.tolist() if ln is not None and not ln.is_cuda else None)
            host = tuple(int(v) for v in (host.tolist() if torch.is_tensor(host) else host)) if host is not None else None
            if not self._want_packed(host, x.shape[0], x.shape[2], self.output_frames(x.shape[2])):
                host = None
            key = self.graph_key(x.shape[0], x.shape[2], in_dtype, ln is not None, torch.float32, device.index, host)
            if key in owner:  # same shape = same graph and static buffers: keep stream order, detach the earlier result
                k, j = owner[key]
                with torch.cuda.stream(pool[k]):
                    outs[j] = (outs[j][0].clone(), outs[j][1].clone())
            else:
                k = i % n_streams
            stream = pool[k]
            if not any(o[0] == k for o in owner.values()):
                stream.wait_event(fork)
            with torch.cuda.stream(stream):
                outs.append(self.forward(audio_signal=x, length=ln, length_host=hl))
            owner[key] = (k, i)
        for k in {o[0] for o in owner.values()}:
            cur.wait_stream(pool[k])
        for enc_t, enc_len in outs:
            enc_t.record_stream(cur)
            enc_len.record_stream(cur)
        return outs

    # ------------------------------------------------------------------------------------------------ weights
    def mark_weights_dirty(self) -> None:
        """Parameters are copied into the library's own arena; call this after modifying them in place
        (``load_state_dict`` and ``.to()`` / ``.cuda()`` do it automatically)."""
        self._dirty = True
        self._graphs.clear()  # captured graphs read the old weight arena

    def _apply(self, fn, *args, **kwargs):  # .to()/.cuda()/.half(): weights must be re-packed
        self._dirty = True
        self.__dict__.get("_graphs", {}).clear()
        return super()._apply(fn, *args, **kwargs)

    def prepare(self, device: Optional[torch.device] = None) -> None:
        """Creates the device handle and packs the current parameters into the library's arena
        (cfb_set_weight / cfb_finalize_weights).  Called automatically by ``forward`` when parameters changed."""
        lib = _lib.load_library()
        if device is None:
            device = next(self.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ConformerEncoder (B200) needs its parameters on a CUDA device; there is no CPU path")
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is None or self._handle_device != index:
            self._destroy()
            cfg = _lib.CfbConfig()
            for k, v in self._cfg.items():
                setattr(cfg, k, int(v))
            cfg.precision = _lib.CFB_PREC_BF16 if self.precision == "bf16" else _lib.CFB_PREC_FP32_VALIDATE
            handle = ctypes.c_void_p()
            _lib.check(lib.cfb_create(ctypes.byref(cfg), index, ctypes.byref(handle)), None, "cfb_create")
            self._handle, self._handle_device = handle, index
        # cfb_set_weight copies with a blocking cudaMemcpy on the legacy default stream, which does not order itself after
        # torch's non-blocking side streams: make sure whatever produced the parameters (a .to() enqueued on the current
        # stream, the dtype conversions below) has completed before the library reads them
        torch.cuda.current_stream(device).synchronize()
        sd = self.state_dict()
        # sinusoid frequencies computed with the same torch expression as the reference (multi_head_attention.py:238-241)
        sd["pos_enc.div_term"] = torch.exp(
            torch.arange(0, self.d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / self.d_model))
        for key, t in sd.items():
            if key.endswith("num_batches_tracked"):
                continue
            t = t.detach()
            if t.dtype not in _DTYPES:
                t = t.float()
            t = t.contiguous()
            if t.is_cuda:
                torch.cuda.current_stream(t.device).synchronize()  # a conversion just enqueued must have finished
            shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
            _lib.check(lib.cfb_set_weight(self._handle, key.encode(), ctypes.c_void_p(t.data_ptr()), _DTYPES[t.dtype],
                                          shape, t.dim()), self._handle, f"cfb_set_weight({key})")
        _lib.check(lib.cfb_finalize_weights(self._handle), self._handle, "cfb_finalize_weights")
        self._dirty = False

    def _destroy(self):
        if self.__dict__.get("_handle") is not None:
            try:
                _lib.load_library().cfb_destroy(self._handle)
            except Exception:  # pragma: no cover - interpreter shutdown
                pass
            self.__dict__["_handle"] = None  # not through nn.Module.__setattr__: torch may already be torn down at exit

    def __del__(self):
        try:
            self._destroy()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass

    # ------------------------------------------------------------------------------------------------ forward
    def output_frames(self, t: int) -> int:
        for _ in range(2):
            t = (t - 1) // 2 + 1
        return t

    def _ensure_workspace(self, nbytes: int, device) -> torch.Tensor:
        ws = self._workspace
        if ws is None or ws.numel() < nbytes or ws.device != device:
            self._workspace = ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return ws

    def forward(self, audio_signal: torch.Tensor, length: Optional[torch.Tensor] = None,
                out_dtype: torch.dtype = torch.float32, length_host=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """audio_signal (B, feat_in, T) float, length (B,) int or None -> (encoded (B, d_out, T'), encoded_len (B,)
        int32).  ``encoded`` is the transposed view of a contiguous (B, T', d_out) buffer, like the reference's
        (conformer_encoder.py:280).  ``length=None`` means every row is T frames long (conformer_encoder.py:243-246).
        ``length_host`` (not in the reference): the same lengths as a host sequence / CPU tensor.  With it (or with a
        CPU ``length``) a mixed-length batch runs in the packed layout -- same result bit for bit, no arithmetic on
        padding (``self.packed``); without it the dense path runs and nothing synchronises with the host."""
        self.update_max_seq_length(seq_length=audio_signal.size(2), device=audio_signal.device)
        return self.forward_for_export(audio_signal=audio_signal, length=length, out_dtype=out_dtype,
                                       length_host=length_host)

    def packed_rows(self, lengths, t: int) -> int:
        """Token rows of the packed layout for these input lengths (csrc/common.cuh: packed_slot_rows)."""
        rows = 0
        for n in lengths:
            n = min(max(int(n), 0), t)
            t2 = (((n + 1) >> 1) + 1) >> 1
            rows += (t2 + 15 + 7) // 8 * 8
        return rows

    def _want_packed(self, host_lens, b, t, t_out) -> bool:
        if host_lens is None or self.packed is False or self.precision != "bf16" or self.out_proj is not None or b > 2048:
            return False
        if self.packed is True:
            return True
        return self.packed_rows(host_lens, t) <= 0.85 * b * t_out

    @torch.no_grad()
    def forward_for_export(self, audio_signal, length=None, out_dtype: torch.dtype = torch.float32, length_host=None):
        if audio_signal.dim() != 3 or audio_signal.size(1) != self._feat_in:
            raise TypeError(f"audio_signal must be (B, {self._feat_in}, T), got {tuple(audio_signal.shape)}")
        if not audio_signal.is_cuda:
            raise RuntimeError("ConformerEncoder (B200) has no CPU path: audio_signal must be a CUDA tensor")
        if self.training:
            raise RuntimeError("ConformerEncoder (B200) is inference-only: call .eval() / .freeze() first")
        device = audio_signal.device
        if self._handle is None or self._dirty or \
                self._handle_device != (device.index if device.index is not None else torch.cuda.current_device()):
            if next(self.parameters()).device != device:
                raise RuntimeError("parameters and audio_signal are on different devices")
            self.prepare(device)
        lib = _lib.load_library()
        b, _, t = audio_signal.shape
        if b < 1 or t < 1:
            raise ValueError("empty batch")
        if audio_signal.dtype not in (torch.float32, torch.bfloat16):
            audio_signal = audio_signal.float()
        feats = audio_signal.contiguous()
        host_lens = None
        if length is not None:
            if length.dim() != 1 or length.numel() != b:
                raise TypeError(f"length must have shape ({b},), got {tuple(length.shape)}")
            if length_host is not None:
                host_lens = tuple(int(v) for v in (length_host.tolist() if torch.is_tensor(length_host) else length_host))
                if len(host_lens) != b:
                    raise TypeError(f"length_host must have {b} entries")
            elif not length.is_cuda:
                host_lens = tuple(int(v) for v in length.tolist())
            length = length.to(device=device, dtype=torch.int64).contiguous()
        t_out = self.output_frames(t)
        if not self._want_packed(host_lens, b, t, t_out):
            host_lens = None  # dense path
        if self._use_graphs and not self._profiling and not torch.cuda.is_current_stream_capturing():
            return self._forward_graph(feats, length, out_dtype, b, t, t_out, device, host_lens)
        encoded = torch.empty(b, t_out, self._feat_out, dtype=out_dtype, device=device)
        encoded_len = torch.empty(b, dtype=torch.int32, device=device)
        nbytes = ctypes.c_size_t()
        if host_lens is not None:
            hl = (ctypes.c_int64 * b)(*host_lens)
            _lib.check(lib.cfb_packed_workspace_bytes(self._handle, hl, b, t, ctypes.byref(nbytes)), self._handle,
                       "cfb_packed_workspace_bytes")
        else:
            _lib.check(lib.cfb_workspace_bytes(self._handle, b, t, ctypes.byref(nbytes)), self._handle, "cfb_workspace_bytes")
        ws = self._ensure_workspace(nbytes.value + 256, device)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            if host_lens is not None:
                _lib.check(lib.cfb_forward_packed(
                    self._handle, ctypes.c_void_p(feats.data_ptr()), _DTYPES[feats.dtype],
                    ctypes.c_void_p(length.data_ptr()), hl, b, t,
                    ctypes.c_void_p(encoded.data_ptr()), _DTYPES[out_dtype], ctypes.c_void_p(encoded_len.data_ptr()),
                    ctypes.c_void_p(ws_ptr), nbytes.value, ctypes.c_void_p(stream)), self._handle, "cfb_forward_packed")
            else:
                _lib.check(lib.cfb_forward(
                    self._handle, ctypes.c_void_p(feats.data_ptr()), _DTYPES[feats.dtype],
                    ctypes.c_void_p(length.data_ptr()) if length is not None else None, b, t,
                    ctypes.c_void_p(encoded.data_ptr()), _DTYPES[out_dtype], ctypes.c_void_p(encoded_len.data_ptr()),
                    ctypes.c_void_p(ws_ptr), nbytes.value, ctypes.c_void_p(stream)), self._handle, "cfb_forward")
        return encoded.transpose(1, 2), encoded_len

    @staticmethod
    def graph_key(b, t, in_dtype, has_len, out_dtype, device_index, host_lens=None):
        """Identity of a captured forward: a packed graph bakes the slot layout in, so it also depends on the lengths."""
        return (b, t, in_dtype, has_len, out_dtype, device_index, host_lens)

    def _forward_graph(self, feats, length, out_dtype, b, t, t_out, device, host_lens=None):
        key = self.graph_key(b, t, feats.dtype, length is not None, out_dtype, device.index, host_lens)
        entry = self._graphs.get(key)
        if entry is None:
            if getattr(self, "_graph_private_ws", False):
                self._workspace = None  # this shape's graph gets a workspace no other graph touches
            static_feats = torch.empty_like(feats)
            static_len = torch.empty(b, dtype=torch.int64, device=device) if length is not None else None
            static_feats.copy_(feats)
            if length is not None:
                static_len.copy_(length)
            self._use_graphs = False
            try:
                side = torch.cuda.Stream(device=device)
                side.wait_stream(torch.cuda.current_stream(device))
                with torch.cuda.stream(side):  # warm-up outside capture (workspace allocation, lazy module loading)
                    self.forward_for_export(static_feats, static_len, out_dtype, length_host=host_lens)
                torch.cuda.current_stream(device).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    enc_t, enc_len = self.forward_for_export(static_feats, static_len, out_dtype, length_host=host_lens)
            finally:
                self._use_graphs = True
            # the entry keeps the workspace the graph was captured with alive
            entry = (graph, static_feats, static_len, enc_t, enc_len, self._workspace)
            self._graphs[key] = entry
            while len(self._graphs) > getattr(self, "_graph_cap", 4):
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
            entry[1].copy_(feats)
            if length is not None:
                entry[2].copy_(length)
        entry[0].replay()
        return entry[3], entry[4]

    # profiling ---------------------------------------------------------------------------------------------------
    def set_profiling(self, on: bool) -> None:
        """Bracket every kernel of subsequent forwards with CUDA events (cfb_set_profiling)."""
        if self._handle is None:
            self.prepare()
        _lib.check(_lib.load_library().cfb_set_profiling(self._handle, int(on)), self._handle, "cfb_set_profiling")
        self._profiling = bool(on)

    def profile_report(self):
        """{label: (launches, total_ms)} for the forwards since the last report (synchronises)."""
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(_lib.load_library().cfb_profile_report(self._handle, buf, len(buf)), self._handle, "cfb_profile_report")
        out = OrderedDict()
        for line in buf.value.decode().splitlines():
            label, n, ms = line.split("\t")
            out[label] = (int(n), float(ms))
        return out

    # debugging / tests -------------------------------------------------------------------------------------------
    def last_launch_count(self) -> int:
        return _lib.load_library().cfb_last_launch_count(self._handle) if self._handle else 0

    def debug_buffer(self, b: int, t: int, name: str) -> torch.Tensor:
        """Raw bytes of an intermediate of the last (b, t) forward (see cfb_debug_buffer)."""
        off, nbytes = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(_lib.load_library().cfb_debug_buffer(self._handle, b, t, name.encode(), ctypes.byref(off),
                                                        ctypes.byref(nbytes)), self._handle, "cfb_debug_buffer")
        base = (self._workspace.data_ptr() + 255) // 256 * 256 - self._workspace.data_ptr()
        return self._workspace[base + off.value: base + off.value + nbytes.value]
