"""Engine caching for the drop-in facades.

The packed native model is a *derived cache* of the module's parameters and buffers: it
is rebuilt whenever any of them changed (optimizer step, ``load_state_dict``, ``.to()``,
BatchNorm running statistics moving during training) -- detected through the tensors'
version counters and storage pointers -- and never otherwise (SURVEY 7.3 "Mode/caching
correctness").
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .engine import EngineConfig, ScoringEngine


class EngineCacheMixin:
    """Mixed into the facade nn.Modules.  Sub-classes implement ``_sf_config()``."""

    def _sf_fingerprint(self) -> Tuple:
        fp = []
        for t in list(self.parameters()) + list(self.buffers()):
            fp.append((t.data_ptr(), t._version, t.device.index))
        return tuple(fp)

    def _sf_engine(self) -> ScoringEngine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("shopformer_b200: the model must live on a CUDA device for inference (no CPU fallback)")
        fp = self._sf_fingerprint()
        cached = self.__dict__.get("_sf_cache")
        if cached is not None and cached[0] == fp:
            return cached[1]
        if cached is not None:
            cached[1].close()
        eng = ScoringEngine(self._sf_config(), self.state_dict(), dev)
        self.__dict__["_sf_cache"] = (fp, eng)
        return eng

    def _sf_config(self) -> EngineConfig:  # pragma: no cover - interface
        raise NotImplementedError


class LazyOutput(dict):
    """Output dict of the variant-1 facade.  ``gcae_reconstructed`` (the pose decoder output,
    SURVEY row f1 -- not needed for the score) is produced on first access instead of on
    every forward; every other way of looking at the dict materialises it first so that it
    behaves like the plain dict the reference returns."""

    def __init__(self, eager: dict, lazy: dict):
        super().__init__(eager)
        self._lazy = dict(lazy)

    def _materialise(self) -> None:
        while self._lazy:
            k, fn = self._lazy.popitem()
            dict.__setitem__(self, k, fn())

    def __missing__(self, key):
        if key in self._lazy:
            val = self._lazy.pop(key)()
            dict.__setitem__(self, key, val)
            return val
        raise KeyError(key)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def __iter__(self):
        self._materialise()
        return dict.__iter__(self)

    def __len__(self):
        return dict.__len__(self) + len(self._lazy)

    def keys(self):
        self._materialise()
        return dict.keys(self)

    def values(self):
        self._materialise()
        return dict.values(self)

    def items(self):
        self._materialise()
        return dict.items(self)

    def __repr__(self):
        self._materialise()
        return dict.__repr__(self)
