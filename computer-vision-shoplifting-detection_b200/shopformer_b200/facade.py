"""Engine caching for the drop-in facades.

The packed native model is a *derived cache* of the module's parameters and buffers: it
is rebuilt whenever any of them changed (optimizer step, ``load_state_dict``, ``.to()``,
BatchNorm running statistics moving during training) -- detected through the tensors'
version counters and storage pointers -- and never otherwise (SURVEY 7.3 "Mode/caching
correctness").
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .engine import EngineConfig, ScoringEngine

PRECISIONS = ("auto", "tc", "bf16", "fp32")


class EngineCacheMixin:
    """Mixed into the facade nn.Modules.  Sub-classes implement ``_sf_config()``.

    Precision policy of the eval-mode native path (every entry point of the facade and of its sub-modules):
    ``model.sf_precision`` if set, else the environment variable ``SHOPFORMER_B200_PRECISION``, else ``"auto"`` =
    the bf16 tcgen05 kernels (score within 1e-2 of the reference, north-star contract) whenever they cover the
    shape, the fp32 kernels (1e-3 contract) otherwise.  Set ``"fp32"`` to force the precise path."""

    sf_precision: Optional[str] = None

    def _sf_resolve_precision(self) -> str:
        p = self.sf_precision or os.environ.get("SHOPFORMER_B200_PRECISION") or "auto"
        if p not in PRECISIONS:
            raise ValueError(f"sf_precision / SHOPFORMER_B200_PRECISION must be one of {PRECISIONS}, got {p!r}")
        return p

    # The fingerprint walks a cached tensor list (no module traversal on the per-batch path).  The list is rebuilt
    # whenever the module tree may have swapped tensors (``_apply`` = .to()/.cuda()/.half(), ``load_state_dict``) and,
    # as a safety net against parameters replaced by attribute assignment, every 256 calls.
    def _sf_tensors(self) -> List[torch.Tensor]:
        d = self.__dict__
        n = d.get("_sf_calls", 0) + 1
        d["_sf_calls"] = n
        lst = d.get("_sf_tensor_list")
        if lst is None or (n & 255) == 0:
            lst = d["_sf_tensor_list"] = list(self.parameters()) + list(self.buffers())
        return lst

    def _apply(self, fn, *a, **kw):
        self.__dict__.pop("_sf_tensor_list", None)
        self.__dict__.pop("_sf_tok_tensor_list", None)
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self.__dict__.pop("_sf_tensor_list", None)
        self.__dict__.pop("_sf_tok_tensor_list", None)
        return super().load_state_dict(*a, **kw)

    def _sf_fingerprint(self) -> Tuple:
        return tuple((t.data_ptr(), t._version) for t in self._sf_tensors())

    def _sf_engine(self) -> ScoringEngine:
        fp = self._sf_fingerprint()
        cached = self.__dict__.get("_sf_cache")
        if cached is not None and cached[0] == fp:
            return cached[1]
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("shopformer_b200: the model must live on a CUDA device for inference (no CPU fallback)")
        if cached is not None:
            cached[1].close()
        eng = ScoringEngine(self._sf_config(), self.state_dict(), dev)
        self.__dict__["_sf_cache"] = (fp, eng)
        return eng

    def _sf_tok_engine(self) -> ScoringEngine:
        """Packed model used for TOKENIZING ONLY while the facade trains with a frozen tokenizer (shopformer_2 stage 2,
        shopformer_2/models/shopformer.py:94-101): keyed on the encoder's tensors alone, so optimizer steps on the
        transformer do not invalidate it.  (Its transformer weights may be stale; nothing but `tokenize` is called.)"""
        enc = self.gcae.encoder
        d = self.__dict__
        lst = d.get("_sf_tok_tensor_list")
        if lst is None:
            lst = d["_sf_tok_tensor_list"] = list(enc.parameters()) + list(enc.buffers())
        fp = tuple((t.data_ptr(), t._version) for t in lst)
        cached = d.get("_sf_tok_cache")
        if cached is not None and cached[0] == fp:
            return cached[1]
        dev = lst[0].device
        if dev.type != "cuda":
            raise RuntimeError("shopformer_b200: the model must live on a CUDA device for inference (no CPU fallback)")
        if cached is not None:
            cached[1].close()
        eng = ScoringEngine(self._sf_config(), self.state_dict(), dev)
        d["_sf_tok_cache"] = (fp, eng)
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

    def copy(self):
        self._materialise()
        return dict(self)

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
