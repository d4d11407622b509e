"""Loads the UNMODIFIED reference ConformerEncoder -- TEST INFRASTRUCTURE ONLY.

Used in the build container to (a) generate the committed golden vectors (tests/golden/make_golden.py) and
(b) cross-check oracle/conformer_oracle.py live when the reference tree is present; on the GPU box, where
/root/reference does not exist, the per-pod install under baseline/_ref (oracle/install_reference.py) lets
bench.py's reference arm / cpu_baseline leg time the reference's own forward.  Nothing in the product imports it.

``import nemo.collections.asr`` fails here (hydra / sox / pytorch_lightning are not installed), so the five
hot-path source files are imported directly after registering stub parent packages and no-op stand-ins for the
three framework symbols they pull in (typecheck, Exportable, NeuralModule) -- SURVEY.md section 8(c).
No reference source is copied: the files are executed from where they lie.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_MARKER = "nemo/collections/asr/modules/conformer_encoder.py"
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference_root() -> str:
    """Search order (GPU-box caveat): $CONFORMER_REF, the mounted tree, then the per-pod installs that
    travel with the repository snapshot (baseline/_ref is written by oracle/install_reference.py, never committed)."""
    cands = [os.environ.get("CONFORMER_REF"), "/root/reference", os.path.join(_REPO, "baseline", "_ref"),
             os.path.join(_REPO, "oracle", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, _MARKER)):
            return c
    return os.environ.get("CONFORMER_REF") or "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _MARKER))


def _stub_package(name: str, path: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod
    return mod


def load_reference_encoder_class():
    """Returns the reference ``ConformerEncoder`` class (conformer_encoder.py:33)."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    if "nemo.collections.asr.modules.conformer_encoder" in sys.modules:
        return sys.modules["nemo.collections.asr.modules.conformer_encoder"].ConformerEncoder
    import torch

    root = REFERENCE_ROOT
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ("nemo", "nemo.collections", "nemo.collections.asr", "nemo.collections.asr.modules",
                 "nemo.collections.asr.parts", "nemo.collections.asr.parts.submodules",
                 "nemo.collections.asr.parts.utils", "nemo.core", "nemo.core.classes"):
        _stub_package(name, os.path.join(root, *name.split(".")))
    # nemo.utils is only needed for `logging` inside neural_types; give it a minimal stand-in
    utils = _stub_package("nemo.utils", os.path.join(root, "nemo/utils"))
    import logging as _logging

    utils.logging = _logging.getLogger("nemo_stub")

    common = types.ModuleType("nemo.core.classes.common")

    def typecheck(*_a, **_k):
        def deco(fn):
            return fn

        return deco

    common.typecheck = typecheck
    sys.modules["nemo.core.classes.common"] = common
    exportable = types.ModuleType("nemo.core.classes.exportable")
    exportable.Exportable = type("Exportable", (), {})
    sys.modules["nemo.core.classes.exportable"] = exportable
    module = types.ModuleType("nemo.core.classes.module")
    module.NeuralModule = torch.nn.Module
    sys.modules["nemo.core.classes.module"] = module
    try:
        importlib.import_module("nemo.core.neural_types")
    except Exception:  # pragma: no cover - fall back to inert placeholders
        nt = types.ModuleType("nemo.core.neural_types")

        class _Any:
            def __init__(self, *a, **k):
                pass

        for n in ("AcousticEncodedRepresentation", "LengthsType", "NeuralType", "SpectrogramType"):
            setattr(nt, n, _Any)
        sys.modules["nemo.core.neural_types"] = nt
    mod = importlib.import_module("nemo.collections.asr.modules.conformer_encoder")
    return mod.ConformerEncoder


def build_reference_encoder(cfg, state_dict=None):
    """Instantiate the reference encoder for an ``oracle.conformer_oracle.EncoderConfig`` and optionally load weights."""
    cls = load_reference_encoder_class()
    enc = cls(
        feat_in=cfg.feat_in, n_layers=cfg.n_layers, d_model=cfg.d_model, feat_out=cfg.feat_out,
        subsampling="striding", subsampling_factor=cfg.subsampling_factor,
        subsampling_conv_channels=cfg.subsampling_conv_channels, ff_expansion_factor=cfg.ff_expansion_factor,
        self_attention_model="rel_pos", n_heads=cfg.n_heads, xscaling=cfg.xscaling,
        conv_kernel_size=cfg.conv_kernel_size, dropout=0.1, dropout_emb=0.0, dropout_att=0.1,
        untie_biases=getattr(cfg, "untie_biases", True),
    )
    if state_dict is not None:
        missing, unexpected = enc.load_state_dict(state_dict, strict=False)
        missing = [m for m in missing if "num_batches_tracked" not in m]
        assert not missing and not unexpected, (missing, unexpected)
    return enc.eval()


def load_reference_ctc_decoder_class():
    """Returns the reference ``ConvASRDecoder`` class (modules/conv_asr.py:397), loaded unmodified.  conv_asr.py
    imports omegaconf (absent here) only for type names; an inert stand-in is registered for it."""
    load_reference_encoder_class()  # registers the stub packages
    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")
        om.MISSING = "???"
        om.ListConfig = list
        om.DictConfig = dict
        om.OmegaConf = type("OmegaConf", (), {})
        sys.modules["omegaconf"] = om
    mod = importlib.import_module("nemo.collections.asr.modules.conv_asr")
    return mod.ConvASRDecoder


def load_reference_filterbank_class():
    """Returns the reference ``FilterbankFeatures`` class (parts/preprocessing/features.py:196), executed unmodified.
    features.py imports librosa, torch_stft and two sibling modules (perturb, segment) that need soundfile / sox: inert
    stand-ins are registered for them.  The only one that contributes arithmetic is ``librosa.filters.mel``, which is
    served by oracle.frontend_oracle.slaney_mel_filters (see that module's header)."""
    load_reference_encoder_class()  # registers the stub packages and nemo.utils.logging
    if "nemo.collections.asr.parts.preprocessing.features" in sys.modules:
        return sys.modules["nemo.collections.asr.parts.preprocessing.features"].FilterbankFeatures
    import numpy as np

    from oracle.frontend_oracle import slaney_mel_filters

    _stub_package("nemo.collections.asr.parts.preprocessing",
                  os.path.join(REFERENCE_ROOT, "nemo/collections/asr/parts/preprocessing"))
    librosa = types.ModuleType("librosa")
    librosa.filters = types.ModuleType("librosa.filters")
    librosa.filters.mel = lambda sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **_k: slaney_mel_filters(sr, n_fft, n_mels, fmin, fmax)
    librosa.util = types.ModuleType("librosa.util")
    librosa.util.tiny = lambda x: np.finfo(np.float32).tiny
    sys.modules.update({"librosa": librosa, "librosa.filters": librosa.filters, "librosa.util": librosa.util})
    ts = types.ModuleType("torch_stft")
    ts.STFT = type("STFT", (), {})
    sys.modules["torch_stft"] = ts
    for name, attr in (("perturb", "AudioAugmentor"), ("segment", "AudioSegment")):
        m = types.ModuleType("nemo.collections.asr.parts.preprocessing." + name)
        setattr(m, attr, type(attr, (), {}))
        sys.modules[m.__name__] = m
    utils = sys.modules["nemo.utils"]
    if not hasattr(utils.logging, "warning"):
        import logging as _logging

        utils.logging = _logging.getLogger("nemo_stub")
    mod = importlib.import_module("nemo.collections.asr.parts.preprocessing.features")
    return mod.FilterbankFeatures


def load_reference_rnnt_classes():
    """Returns the reference ``(RNNTDecoder, RNNTJoint, GreedyBatchedRNNTInfer)`` classes (modules/rnnt.py:51,613;
    parts/submodules/rnnt_greedy_decoding.py:358), executed unmodified.  The files import framework symbols that are
    absent here: ``nemo.core.NeuralModule`` (stand-in: torch.nn.Module + the ``as_frozen`` context manager of
    nemo/core/classes/module.py:67-84), ``nemo.core.classes.Typing`` (empty class) and ``typecheck`` (no-op)."""
    load_reference_encoder_class()  # registers the stub packages and nemo.utils.logging
    name = "nemo.collections.asr.parts.submodules.rnnt_greedy_decoding"
    if name in sys.modules:
        g = sys.modules[name]
        r = sys.modules["nemo.collections.asr.modules.rnnt"]
        return r.RNNTDecoder, r.RNNTJoint, g.GreedyBatchedRNNTInfer
    import contextlib

    import torch

    class NeuralModule(torch.nn.Module):
        def freeze(self):
            for p in self.parameters():
                p.requires_grad = False
            self.eval()

        def unfreeze(self):
            for p in self.parameters():
                p.requires_grad = True
            self.train()

        @contextlib.contextmanager
        def as_frozen(self):
            training = self.training
            flags = {n: p.requires_grad for n, p in self.named_parameters()}
            self.freeze()
            try:
                yield
            finally:
                for n, p in self.named_parameters():
                    p.requires_grad = flags[n]
                self.train(training)

    sys.modules["nemo.core"].NeuralModule = NeuralModule
    classes = sys.modules["nemo.core.classes"]
    classes.typecheck = sys.modules["nemo.core.classes.common"].typecheck
    classes.Typing = type("Typing", (), {})
    for pkg in ("nemo.collections.common", "nemo.collections.common.parts"):
        _stub_package(pkg, os.path.join(REFERENCE_ROOT, *pkg.split(".")))
    r = importlib.import_module("nemo.collections.asr.modules.rnnt")
    g = importlib.import_module(name)
    return r.RNNTDecoder, r.RNNTJoint, g.GreedyBatchedRNNTInfer


def load_reference_collation():
    """Returns ``(audio_to_text module, collections module)`` of the reference, executed unmodified:
    ``data/audio_to_text.py`` (``_speech_collate_fn`` :48-99, ``BucketingIterator`` :1515-1533) and
    ``common/parts/preprocessing/collections.py`` (``ASRAudioText`` manifest filtering :88-216).  Their imports that are
    absent here (braceexpand, webdataset, frozendict, the text cleaners, tokenizers, nemo.core Dataset classes,
    ``deprecated``) get inert stand-ins; none of them contributes to collation or manifest filtering."""
    load_reference_filterbank_class()  # stub packages, nemo.utils.logging, librosa stand-in, features.py
    name = "nemo.collections.asr.data.audio_to_text"
    if name in sys.modules:
        return sys.modules[name], sys.modules["nemo.collections.common.parts.preprocessing.collections"]
    import torch

    for pkg in ("nemo.collections.common", "nemo.collections.common.parts", "nemo.collections.common.parts.preprocessing",
                "nemo.collections.asr.data"):
        if pkg not in sys.modules:
            _stub_package(pkg, os.path.join(REFERENCE_ROOT, *pkg.split(".")))
    for mod_name in ("braceexpand", "webdataset"):
        if mod_name not in sys.modules:
            sys.modules[mod_name] = types.ModuleType(mod_name)
    if "frozendict" not in sys.modules:
        fd = types.ModuleType("frozendict")
        fd.frozendict = dict
        sys.modules["frozendict"] = fd
    cleaners = types.ModuleType("nemo.collections.common.parts.preprocessing.cleaners")
    cleaners.clean_text = lambda s, *a, **k: s
    sys.modules[cleaners.__name__] = cleaners
    tok = types.ModuleType("nemo.collections.common.tokenizers")
    tok.TokenizerSpec = type("TokenizerSpec", (), {})
    tok.AggregateTokenizer = type("AggregateTokenizer", (), {})
    sys.modules[tok.__name__] = tok
    sys.modules["nemo.collections.common"].tokenizers = tok
    classes = sys.modules["nemo.core.classes"]
    classes.Dataset = torch.utils.data.Dataset
    classes.IterableDataset = torch.utils.data.IterableDataset
    if not hasattr(classes, "typecheck"):
        classes.typecheck = sys.modules["nemo.core.classes.common"].typecheck
    dec = types.ModuleType("nemo.utils.decorators")

    def deprecated(*_a, **_k):
        return lambda fn: fn

    dec.deprecated = deprecated
    sys.modules[dec.__name__] = dec
    utils = sys.modules["nemo.utils"]
    if not hasattr(utils.logging, "info"):
        import logging as _logging

        utils.logging = _logging.getLogger("nemo_stub")
    coll = importlib.import_module("nemo.collections.common.parts.preprocessing.collections")
    a2t = importlib.import_module(name)
    return a2t, coll


def load_reference_sample_greedy_class():
    """The reference's sample-level ``GreedyRNNTInfer`` (rnnt_greedy_decoding.py:191), from the same unmodified module."""
    load_reference_rnnt_classes()
    return sys.modules["nemo.collections.asr.parts.submodules.rnnt_greedy_decoding"].GreedyRNNTInfer


def load_reference_wer_class():
    """Returns the reference ``WER`` metric class (metrics/wer.py:67), executed unmodified; its greedy CTC collapse is
    ``WER.ctc_decoder_predictions_tensor`` (:122-188).  wer.py imports ``editdistance`` and ``torchmetrics`` (both absent
    here): ``editdistance`` gets an inert stand-in (the collapse never calls it) and ``torchmetrics.Metric`` a minimal base
    class that accepts the constructor keywords and ``add_state`` calls WER makes."""
    load_reference_rnnt_classes()  # stub packages, nemo.utils.logging, rnnt_utils.Hypothesis
    name = "nemo.collections.asr.metrics.wer"
    if name in sys.modules:
        return sys.modules[name].WER
    if "editdistance" not in sys.modules:
        ed = types.ModuleType("editdistance")
        ed.eval = lambda a, b: (_ for _ in ()).throw(NotImplementedError("editdistance stand-in"))
        sys.modules["editdistance"] = ed
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class Metric:
            def __init__(self, *args, **kwargs):
                pass

            def add_state(self, name, default, dist_reduce_fx=None, persistent=False):
                setattr(self, name, default)

        tm.Metric = Metric
        sys.modules["torchmetrics"] = tm
    _stub_package("nemo.collections.asr.metrics", os.path.join(REFERENCE_ROOT, "nemo/collections/asr/metrics"))
    utils = sys.modules["nemo.utils"]
    if not hasattr(utils.logging, "info"):
        import logging as _logging

        utils.logging = _logging.getLogger("nemo_stub")
    return importlib.import_module(name).WER


def reference_ctc_collapse(predictions, lengths, blank_id: int, fold_consecutive: bool = True):
    """Token ids per utterance as the reference's ``WER.ctc_decoder_predictions_tensor`` produces them.  The method returns
    text, so the vocabulary handed to WER is one distinct private-use character per class: text <-> ids is a bijection."""
    import torch

    wer = load_reference_wer_class()(vocabulary=[chr(0xE000 + i) for i in range(blank_id)], fold_consecutive=fold_consecutive,
                                     log_prediction=False)
    lens = None if lengths is None else torch.as_tensor(lengths)
    texts = wer.ctc_decoder_predictions_tensor(torch.as_tensor(predictions), lens)
    return [[ord(c) - 0xE000 for c in t] for t in texts]
