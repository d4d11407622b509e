"""Selection surface: the ``model.encoder`` block of the reference's Hydra recipes (configs/conformer_*.yaml).

Hydra / OmegaConf are not dependencies here; this is the small subset the encoder block needs: YAML loading,
``${a.b.c}`` interpolation against the whole document, and ``_target_`` resolution
(nemo/core/classes/common.py:426-444 -> hydra.utils.instantiate).  Pointing ``_target_`` at
``conformer_nemo_b200.ConformerEncoder`` (or leaving the reference's class path, which is mapped to it) builds the
B200 encoder with the same keyword arguments.
"""
from __future__ import annotations

import importlib
import re
from typing import Any, Dict

import yaml

REFERENCE_TARGETS = {
    "nemo.collections.asr.modules.ConformerEncoder",
    "nemo.collections.asr.modules.conformer_encoder.ConformerEncoder",
}
_INTERP = re.compile(r"\$\{([^}]+)\}")


def _lookup(root: Dict[str, Any], dotted: str):
    node: Any = root
    for part in dotted.split("."):
        if not isinstance(node, dict) or part not in node:
            raise KeyError(f"interpolation key '{dotted}' not found")
        node = node[part]
    return node


def _resolve(value, root, depth=0):
    if depth > 16:
        raise ValueError("interpolation cycle")
    if isinstance(value, dict):
        return {k: _resolve(v, root, depth) for k, v in value.items()}
    if isinstance(value, list):
        return [_resolve(v, root, depth) for v in value]
    if isinstance(value, str):
        m = _INTERP.fullmatch(value.strip())
        if m:  # whole-value interpolation keeps the referenced type (e.g. feat_in: ${model.preprocessor.features})
            return _resolve(_lookup(root, m.group(1)), root, depth + 1)
        if _INTERP.search(value):
            return _INTERP.sub(lambda mm: str(_resolve(_lookup(root, mm.group(1)), root, depth + 1)), value)
    return value


def load_encoder_config(path_or_dict, overrides: Dict[str, Any] = None) -> Dict[str, Any]:
    """Returns the resolved ``model.encoder`` block of a recipe (file path or already-parsed dict)."""
    if isinstance(path_or_dict, dict):
        doc = path_or_dict
    else:
        with open(path_or_dict) as f:
            doc = yaml.safe_load(f)
    enc = doc["model"]["encoder"] if "model" in doc else doc.get("encoder", doc)
    enc = _resolve(dict(enc), doc)
    if overrides:
        enc.update(overrides)
    return enc


def instantiate_encoder(encoder_cfg: Dict[str, Any], **extra):
    """``from_config_dict`` for the encoder block: builds ``_target_(**kwargs)``.  The reference's own class path is
    mapped to the B200 encoder, so an untouched recipe works too."""
    cfg = dict(encoder_cfg)
    target = cfg.pop("_target_", "conformer_nemo_b200.ConformerEncoder")
    cfg.update(extra)
    if target in REFERENCE_TARGETS:
        from .encoder import ConformerEncoder

        return ConformerEncoder(**cfg)
    module_name, _, cls_name = target.rpartition(".")
    cls = getattr(importlib.import_module(module_name), cls_name)
    return cls(**cfg)
