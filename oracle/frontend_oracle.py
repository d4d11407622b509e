"""CPU oracle of the log-mel front-end -- TEST INFRASTRUCTURE ONLY.

A torch-fp32 restatement of ``FilterbankFeatures.forward`` in eval mode
(nemo/collections/asr/parts/preprocessing/features.py:358-453) for the configuration every Conformer recipe uses
(``AudioToMelSpectrogramPreprocessor``: hann window, ``normalize: per_feature``, ``log: true``, guard "add", power 2,
no frame splicing, ``exact_pad: false``).  Imported only by tests/, __graft_entry__.smoke() and tools/; the product
path is the CUDA kernel behind ``cfb_op_logmel``.

Pin: the reference class itself is executed in the build container through oracle/reference_loader.py and its outputs
are committed as tests/golden/frontend_*.npz (tests/golden/make_golden_frontend.py).  The one ingredient the reference
does not contain is the mel filter bank, which it takes from ``librosa.filters.mel`` (features.py:306-309; librosa is
not installed here): ``slaney_mel_filters`` restates librosa's published algorithm (Slaney-style mel scale, htk=False,
norm="slaney") and is handed to the reference through a stand-in ``librosa`` module, so the pin covers everything
except those filter values.  At run time the filters come from the checkpoint's ``featurizer.fb`` buffer anyway.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_filters(sr=16000, n_fft=512, n_mels=80, fmin=0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults htk=False, norm="slaney": (n_mels, 1 + n_fft/2)
    float32 triangular filters on the Slaney mel scale, each scaled to unit area."""
    fmax = float(sr) / 2 if fmax is None else float(fmax)
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None].astype(np.float32) if False else enorm[:, None]
    return w.astype(np.float32)


def seq_len_frames(lengths: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """features.py:347-353 with center=True: pad_amount = n_fft // 2 * 2; floor((len + pad - n_fft) / hop) + 1 in
    float32, then int64."""
    pad_amount = n_fft // 2 * 2
    return (torch.floor((lengths.float() + pad_amount - n_fft) / hop) + 1).to(torch.long)


def filterbank_features(x: torch.Tensor, lengths: torch.Tensor, window: torch.Tensor, fb: torch.Tensor, *, n_fft=512,
                        hop=160, preemph=0.97, log_guard=2.0 ** -24, pad_to=0, std_eps=1e-5):
    """x (B, L) fp32 waveforms, lengths (B,) samples.  Returns (features (B, n_mels, T_out) fp32, seq_len (B,) int64)
    exactly as features.py:358-453 does in eval mode for the supported configuration."""
    win_length = window.numel()
    seq_len = seq_len_frames(lengths, n_fft, hop)                                              # :359
    if preemph is not None:                                                                    # :371-372
        x = torch.cat((x[:, 0].unsqueeze(1), x[:, 1:] - preemph * x[:, :-1]), dim=1)
    spec = torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=win_length, center=True,     # :292-300, :376
                      window=window.float(), return_complex=True)
    spec = torch.view_as_real(spec)
    mag = torch.sqrt(spec.pow(2).sum(-1))                                                      # :383-385
    power = mag.pow(2.0)                                                                       # :393-394
    mel = torch.matmul(fb.to(power.dtype), power)                                              # :397
    feat = torch.log(mel + log_guard)                                                          # :400-402
    # normalize_batch(..., "per_feature") :54-70: statistics over the valid frames of every (utterance, mel bin);
    # torch.std is the unbiased estimator; CONSTANT = 1e-5 is added to the std
    mean = torch.zeros(feat.shape[0], feat.shape[1])
    std = torch.zeros(feat.shape[0], feat.shape[1])
    for i in range(feat.shape[0]):
        n = int(seq_len[i])
        if n == 1:
            raise ValueError("normalize_batch with `per_feature` received a tensor of length 1")
        mean[i] = feat[i, :, :n].mean(dim=1)
        std[i] = feat[i, :, :n].std(dim=1)
    std = std + std_eps
    feat = (feat - mean.unsqueeze(2)) / std.unsqueeze(2)
    t = feat.shape[-1]                                                                         # :436-443
    mask = torch.arange(t)[None, :] >= seq_len[:, None]
    feat = feat.masked_fill(mask.unsqueeze(1), 0.0)
    if pad_to and pad_to > 0 and t % pad_to:                                                   # :447-451
        feat = torch.nn.functional.pad(feat, (0, pad_to - t % pad_to), value=0.0)
    return feat, seq_len


def synthetic_waveforms(batch: int, samples: int, lengths, seed: int):
    """Speech-like test signals: a few drifting harmonics + noise, with silence after each utterance's length."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(samples, dtype=torch.float32) / 16000.0
    x = torch.zeros(batch, samples)
    for b in range(batch):
        f0 = 90.0 + 140.0 * float(torch.rand((), generator=g))
        for h in range(1, 9):
            amp = float(torch.rand((), generator=g)) / h
            x[b] += amp * torch.sin(2 * math.pi * (f0 * h) * t * (1.0 + 0.05 * torch.sin(2 * math.pi * 1.3 * t)))
        x[b] += 0.05 * torch.randn(samples, generator=g)
        x[b] *= 0.1
        x[b, int(lengths[b]):] = 0.0
    return x, torch.tensor(list(lengths), dtype=torch.int64)
