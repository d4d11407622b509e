"""CPU oracle for the Conformer encoder forward pass -- TEST INFRASTRUCTURE ONLY.

This file is a from-scratch, *functional* restatement (plain torch fp32 ops over a flat
``state_dict``) of the algorithm the reference implements in

    nemo/collections/asr/modules/conformer_encoder.py:231-281   (ConformerEncoder.forward)
    nemo/collections/asr/parts/submodules/subsampling.py:163-176,272-282
    nemo/collections/asr/parts/submodules/multi_head_attention.py:69-115,159-210,235-316
    nemo/collections/asr/parts/submodules/conformer_modules.py:88-121,160-180,195-200

It is the checker for the CUDA path, never the thing shipped or measured: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  The product package must never import anything under ``oracle/``.

Parity pin: the reference ships no golden vectors for this path, so the
oracle is pinned against outputs of the *reference itself*, generated in the build container by
``tests/golden/make_golden.py`` (which imports the unmodified reference sources through
``oracle/reference_loader.py``) and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks oracle == golden to fp32 round-off on every run.

The oracle deliberately does not share structure with the reference: no nn.Module tree, rel_shift is
an index gather (proved bit-identical to the pad/view trick by ``tests/test_oracle_golden.py``), masks
are built from lengths on the fly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class EncoderConfig:
    """Constructor surface of the reference encoder that the oracle understands
    (conformer_encoder.py:111-132).  Dropouts are identity in eval and are not represented."""

    feat_in: int = 80
    n_layers: int = 17
    d_model: int = 512
    feat_out: int = -1
    subsampling_factor: int = 4
    subsampling_conv_channels: int = -1
    ff_expansion_factor: int = 4
    n_heads: int = 8
    xscaling: bool = True
    conv_kernel_size: int = 31
    untie_biases: bool = True  # False: one (pos_bias_u, pos_bias_v) pair shared by every layer (conformer_encoder.py:165-173)

    @property
    def conv_channels(self) -> int:
        return self.d_model if self.subsampling_conv_channels == -1 else self.subsampling_conv_channels

    @property
    def n_sub_stages(self) -> int:
        return int(math.log(self.subsampling_factor, 2))


# --------------------------------------------------------------------------------------------
# lengths
# --------------------------------------------------------------------------------------------
def subsampled_lengths(lengths: Tensor, n_stages: int) -> Tensor:
    """subsampling.py:272-282 with padding=1, kernel=3, stride=2, ceil_mode=False: the length is
    pushed through float32 arithmetic and truncated to int32."""
    v = lengths
    for _ in range(n_stages):
        v = torch.floor((v.to(torch.float32) + (2.0 * 1 - 3)) / 2.0 + 1.0)
    return v.to(torch.int32)


def subsampled_extent(t: int, n_stages: int) -> int:
    """Time extent of the padded batch after the strided convolutions (Conv2d k=3,s=2,p=1)."""
    for _ in range(n_stages):
        t = (t + 2 - 3) // 2 + 1
    return t


# --------------------------------------------------------------------------------------------
# positional table
# --------------------------------------------------------------------------------------------
def rel_pos_table(t_out: int, d_model: int, device=None) -> Tensor:
    """multi_head_attention.py:235-248,285-316.  Returns pos_emb of shape (2*t_out-1, d_model) whose row
    k encodes relative position (t_out-1-k): sin in even columns, cos in odd columns.  The reference
    builds a table for a larger max length and slices the centred 2*t_out-1 rows; the values of the
    slice do not depend on the table length, so the slice is generated directly."""
    pos = torch.arange(t_out - 1, -t_out, -1, dtype=torch.float32, device=device).unsqueeze(1)
    div = torch.exp(
        torch.arange(0, d_model, 2, dtype=torch.float32, device=device) * -(math.log(10000.0) / d_model)
    )
    table = torch.zeros(2 * t_out - 1, d_model, dtype=torch.float32, device=device)
    table[:, 0::2] = torch.sin(pos * div)
    table[:, 1::2] = torch.cos(pos * div)
    return table


def rel_shift_gather(bd_full: Tensor) -> Tensor:
    """multi_head_attention.py:159-170 followed by the ``[..., :T]`` slice of :206, written as the gather
    out[..., i, j] = bd_full[..., i, T-1+j-i]."""
    t = bd_full.shape[-2]
    i = torch.arange(t, device=bd_full.device).unsqueeze(1)
    j = torch.arange(t, device=bd_full.device).unsqueeze(0)
    idx = (t - 1 + j - i).expand(*bd_full.shape[:-2], t, t)
    return torch.gather(bd_full, -1, idx)


# --------------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------------
def _subsample(sd: Dict[str, Tensor], cfg: EncoderConfig, feats_btf: Tensor) -> Tensor:
    """subsampling.py:172-175 (striding branch :99-116): n_stages x [Conv2d 3x3 s2 p1 + ReLU], then
    (b,c,t,f)->(b,t,c*f) and the output Linear."""
    y = feats_btf.unsqueeze(1)
    for s in range(cfg.n_sub_stages):
        y = F.relu(F.conv2d(y, sd[f"pre_encode.conv.{2 * s}.weight"], sd[f"pre_encode.conv.{2 * s}.bias"],
                            stride=2, padding=1))
    b, c, t, f = y.shape
    y = y.permute(0, 2, 1, 3).reshape(b, t, c * f)
    return F.linear(y, sd["pre_encode.out.weight"], sd["pre_encode.out.bias"])


def _layer_norm(sd, prefix: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _feed_forward(sd, prefix: str, x: Tensor) -> Tensor:
    """conformer_modules.py:195-200."""
    h = F.silu(F.linear(x, sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"]))
    return F.linear(h, sd[prefix + ".linear2.weight"], sd[prefix + ".linear2.bias"])


def _rel_pos_attention(sd, prefix: str, x: Tensor, pos_emb: Tensor, valid: Tensor, n_heads: int,
                       stages: Optional[dict] = None) -> Tensor:
    """multi_head_attention.py:172-210 + :94-115.  ``valid`` is (B,T) bool; the reference's attention mask is
    ~(valid_q & valid_k) (conformer_encoder.py:260-267)."""
    b, t, d = x.shape
    dk = d // n_heads
    q = F.linear(x, sd[prefix + ".linear_q.weight"], sd[prefix + ".linear_q.bias"]).view(b, t, n_heads, dk)
    k = F.linear(x, sd[prefix + ".linear_k.weight"], sd[prefix + ".linear_k.bias"]).view(b, t, n_heads, dk)
    v = F.linear(x, sd[prefix + ".linear_v.weight"], sd[prefix + ".linear_v.bias"]).view(b, t, n_heads, dk)
    p = F.linear(pos_emb, sd[prefix + ".linear_pos.weight"]).view(2 * t - 1, n_heads, dk)
    qu = (q + sd[prefix + ".pos_bias_u"]).permute(0, 2, 1, 3)  # (b,h,t,dk)
    qv = (q + sd[prefix + ".pos_bias_v"]).permute(0, 2, 1, 3)
    kh = k.permute(0, 2, 3, 1)  # (b,h,dk,t)
    vh = v.permute(0, 2, 1, 3)  # (b,h,t,dk)
    ph = p.permute(1, 2, 0)  # (h,dk,2t-1)
    ac = torch.matmul(qu, kh)
    bd = rel_shift_gather(torch.matmul(qv, ph))
    scores = (ac + bd) / math.sqrt(dk)
    masked = ~(valid.unsqueeze(2) & valid.unsqueeze(1)).unsqueeze(1)  # (b,1,t,t) True = excluded
    scores = scores.masked_fill(masked, -10000.0)
    attn = torch.softmax(scores, dim=-1).masked_fill(masked, 0.0)
    ctx = torch.matmul(attn, vh).permute(0, 2, 1, 3).reshape(b, t, d)
    if stages is not None:
        stages["ctx"] = ctx
    return F.linear(ctx, sd[prefix + ".linear_out.weight"], sd[prefix + ".linear_out.bias"])


def _conv_module(sd, prefix: str, x: Tensor, valid: Tensor, kernel_size: int,
                 stages: Optional[dict] = None) -> Tensor:
    """conformer_modules.py:160-180 with eval-mode BatchNorm1d."""
    d = x.shape[-1]
    y = x.transpose(1, 2)
    y = F.conv1d(y, sd[prefix + ".pointwise_conv1.weight"], sd[prefix + ".pointwise_conv1.bias"])
    y = F.glu(y, dim=1)
    y = y.masked_fill(~valid.unsqueeze(1), 0.0)
    if stages is not None:
        stages["glu"] = y.transpose(1, 2)
    y = F.conv1d(y, sd[prefix + ".depthwise_conv.weight"], sd[prefix + ".depthwise_conv.bias"],
                 padding=(kernel_size - 1) // 2, groups=d)
    y = F.batch_norm(y, sd[prefix + ".batch_norm.running_mean"], sd[prefix + ".batch_norm.running_var"],
                     sd[prefix + ".batch_norm.weight"], sd[prefix + ".batch_norm.bias"], False, 0.0, 1e-5)
    y = F.silu(y)
    if stages is not None:
        stages["dw"] = y.transpose(1, 2)
    y = F.conv1d(y, sd[prefix + ".pointwise_conv2.weight"], sd[prefix + ".pointwise_conv2.bias"])
    return y.transpose(1, 2)


def _layer(sd, i: int, cfg: EncoderConfig, x: Tensor, pos_emb: Tensor, valid: Tensor,
           stages: Optional[dict] = None) -> Tensor:
    """conformer_modules.py:88-121."""
    p = f"layers.{i}"
    r = x + 0.5 * _feed_forward(sd, p + ".feed_forward1", _layer_norm(sd, p + ".norm_feed_forward1", x))
    if stages is not None:
        stages["ff1"] = r
    r = r + _rel_pos_attention(sd, p + ".self_attn", _layer_norm(sd, p + ".norm_self_att", r), pos_emb, valid,
                               cfg.n_heads, stages)
    if stages is not None:
        stages["att"] = r
    r = r + _conv_module(sd, p + ".conv", _layer_norm(sd, p + ".norm_conv", r), valid, cfg.conv_kernel_size, stages)
    if stages is not None:
        stages["conv"] = r
    r = r + 0.5 * _feed_forward(sd, p + ".feed_forward2", _layer_norm(sd, p + ".norm_feed_forward2", r))
    return _layer_norm(sd, p + ".norm_out", r)


# --------------------------------------------------------------------------------------------
# entry point
# --------------------------------------------------------------------------------------------
@torch.no_grad()
def encoder_forward(sd: Dict[str, Tensor], cfg: EncoderConfig, audio_signal: Tensor,
                    length: Optional[Tensor] = None, stages: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """ConformerEncoder.forward(audio_signal (B,feat_in,T), length (B,)) -> (encoded (B,d_out,T'), encoded_len
    (B,) int32), conformer_encoder.py:231-281.  ``stages`` (optional dict) receives per-stage intermediates
    of layer 0 for kernel-level tests."""
    b, _, t = audio_signal.shape
    if length is None:  # conformer_encoder.py:243-246
        length = torch.full((b,), t, dtype=torch.int32, device=audio_signal.device)
    enc_len = subsampled_lengths(length, cfg.n_sub_stages)
    x = _subsample(sd, cfg, audio_signal.transpose(1, 2).to(torch.float32))
    if cfg.xscaling:  # multi_head_attention.py:305-306
        x = x * math.sqrt(cfg.d_model)
    t_out = x.shape[1]
    pos_emb = rel_pos_table(t_out, cfg.d_model, device=x.device)
    valid = torch.arange(t_out, device=x.device).unsqueeze(0) < enc_len.unsqueeze(1)  # conformer_encoder.py:296-299
    if stages is not None:
        stages["pre_encode"] = x
        stages["pos_emb"] = pos_emb
    for i in range(cfg.n_layers):
        x = _layer(sd, i, cfg, x, pos_emb, valid, stages if i == 0 else None)
        if stages is not None and i == 0:
            stages["layer0"] = x
    if "out_proj.weight" in sd:  # conformer_encoder.py:277-278
        x = F.linear(x, sd["out_proj.weight"], sd["out_proj.bias"])
    return x.transpose(1, 2), enc_len


# --------------------------------------------------------------------------------------------
# fixtures: reference-shaped random weights
# --------------------------------------------------------------------------------------------
def expected_state_shapes(cfg: EncoderConfig) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of the reference ``state_dict``; ``num_batches_tracked`` omitted."""
    d, c, ff, h = cfg.d_model, cfg.conv_channels, cfg.d_model * cfg.ff_expansion_factor, cfg.n_heads
    f = cfg.feat_in
    for _ in range(cfg.n_sub_stages):
        f = (f + 2 - 3) // 2 + 1
    shapes: Dict[str, Tuple[int, ...]] = {}
    cin = 1
    for s in range(cfg.n_sub_stages):
        shapes[f"pre_encode.conv.{2 * s}.weight"] = (c, cin, 3, 3)
        shapes[f"pre_encode.conv.{2 * s}.bias"] = (c,)
        cin = c
    shapes["pre_encode.out.weight"] = (d, c * f)
    shapes["pre_encode.out.bias"] = (d,)
    for i in range(cfg.n_layers):
        p = f"layers.{i}."
        for n in ("norm_feed_forward1", "norm_self_att", "norm_conv", "norm_feed_forward2", "norm_out"):
            shapes[p + n + ".weight"] = (d,)
            shapes[p + n + ".bias"] = (d,)
        for n in ("feed_forward1", "feed_forward2"):
            shapes[p + n + ".linear1.weight"] = (ff, d)
            shapes[p + n + ".linear1.bias"] = (ff,)
            shapes[p + n + ".linear2.weight"] = (d, ff)
            shapes[p + n + ".linear2.bias"] = (d,)
        for n in ("linear_q", "linear_k", "linear_v", "linear_out"):
            shapes[p + "self_attn." + n + ".weight"] = (d, d)
            shapes[p + "self_attn." + n + ".bias"] = (d,)
        shapes[p + "self_attn.linear_pos.weight"] = (d, d)
        shapes[p + "self_attn.pos_bias_u"] = (h, d // h)
        shapes[p + "self_attn.pos_bias_v"] = (h, d // h)
        shapes[p + "conv.pointwise_conv1.weight"] = (2 * d, d, 1)
        shapes[p + "conv.pointwise_conv1.bias"] = (2 * d,)
        shapes[p + "conv.depthwise_conv.weight"] = (d, 1, cfg.conv_kernel_size)
        shapes[p + "conv.depthwise_conv.bias"] = (d,)
        for n in ("weight", "bias", "running_mean", "running_var"):
            shapes[p + "conv.batch_norm." + n] = (d,)
        shapes[p + "conv.pointwise_conv2.weight"] = (d, d, 1)
        shapes[p + "conv.pointwise_conv2.bias"] = (d,)
    if cfg.feat_out > 0 and cfg.feat_out != d:
        shapes["out_proj.weight"] = (cfg.feat_out, d)
        shapes[("out_proj.bias")] = (cfg.feat_out,)
    return shapes


def random_state_dict(cfg: EncoderConfig, seed: int = 0) -> Dict[str, Tensor]:
    """Synthetic weights with the statistics of torch's default initialisers (U(-1/sqrt(fan_in), +)), plus the
    fixture randomisation SURVEY.md section 8(c) asks for: pos_bias_u/v ~ N(0, 0.1^2), BN running_mean ~ N(0,0.1^2),
    running_var ~ U(0.75,1.25), LayerNorm/BN affine perturbed around (1, 0).  (The reference's own defaults for
    those are degenerate zeros/ones and would leave the u/v and BN-fold paths untested.)"""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for key, shape in expected_state_shapes(cfg).items():
        if "pos_bias" in key or key.endswith("running_mean"):
            sd[key] = torch.randn(shape, generator=g) * 0.1
        elif key.endswith("running_var"):
            sd[key] = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif ".norm_" in key or "batch_norm" in key:
            base = 1.0 if key.endswith("weight") else 0.0
            sd[key] = base + torch.randn(shape, generator=g) * 0.1
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if key.endswith(".bias"):
                w = expected_state_shapes(cfg)[key[: -len("bias")] + "weight"]
                fan_in = 1
                for s in w[1:]:
                    fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    if not cfg.untie_biases:
        # the shared pair is a local of the reference constructor, not an attribute (conformer_encoder.py:165-173): the
        # state_dict lists it under every layer's name, all with the same values
        for name in ("pos_bias_u", "pos_bias_v"):
            for i in range(1, cfg.n_layers):
                sd[f"layers.{i}.self_attn.{name}"] = sd[f"layers.0.self_attn.{name}"].clone()
    return sd


def synthetic_batch(b: int, feat_in: int, t: int, lengths=None, seed: int = 1234) -> Tuple[Tensor, Tensor]:
    """randn log-mel-like features, frames t >= len zeroed as the preprocessor does."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, feat_in, t, generator=g)
    if lengths is None:
        length = torch.full((b,), t, dtype=torch.int64)
    else:
        length = torch.as_tensor(lengths, dtype=torch.int64)
    x = x * (torch.arange(t).view(1, 1, t) < length.view(b, 1, 1))
    return x, length
