"""Shared test helpers: build drop-in models with the deterministic synthetic weights."""
from __future__ import annotations

import numpy as np
import torch

from shopformer_b200 import configs as CFG
from shopformer_b200.synthetic import synth_state_dict


def build_model(dropin1, dropin2, name: str, seed: int = 0):
    args = CFG.ctor_args(name)
    if CFG.variant_of(name) == 1:
        model = dropin1["models"].Shopformer(**args)
    else:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = dropin2["models"].Shopformer(args)
    model.load_state_dict(synth_state_dict(model.state_dict(), seed=seed), strict=True)
    return model.eval()


def oracle_kwargs(model, name: str):
    enc = model.gcae.encoder
    var = CFG.variant_of(name)
    return dict(variant=var, strides=list(enc.strides), nhead=model.transformer.nhead,
                pool_tokens=(enc.num_tokens if enc._needs_pooling else None))


def rel_err(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))


def max_abs_rel(a, b) -> float:
    """max |a-b| / max |b|  (for tensors with entries near zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))
