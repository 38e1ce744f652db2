"""``torch.ops.shopformer_b200.*`` -- the scoring path as PyTorch custom ops (BASELINE north star / SURVEY 8b).

A thin shim over the C ABI: each op validates its tensors, allocates the output and calls the corresponding
``sf_*`` entry point on the current CUDA stream through :class:`ScoringEngine`.  Registered with ``torch.library`` so
the ops are visible to the dispatcher (profiler ranges, ``torch.compile`` graph capture through the fake
implementations below, ``torch.library.opcheck``).  A packed model travels as an integer handle
(:func:`model_handle`); ops are CUDA-only -- there is no CPU kernel, calling one on CPU tensors raises
``NotImplementedError`` from the dispatcher.

    window_normalize(raw (B,T,K,3), V, normalize)            -> poses (B,2,T,V)      sf_normalize_windows
    tokenize(poses (B,C,T,V), model, precision)              -> tokens (B,S,D)       sf_tokenize
    reconstruct_tokens(tokens (B,S,D), model, precision)     -> recon (B,S,D)        sf_reconstruct_tokens
    normality_score(tokens, recon, model, reduction)         -> scores (B) | (B,S)   sf_normality_score
    score_fused(poses, model, reduction, precision)          -> scores (B) | (B,S)   sf_score_windows
    score_fused_full(poses, model, precision)                -> scores, tokens, recon  sf_score_windows (every output)

Reference methods replaced: ``GCAEEncoder.forward`` (shopformer/models/gcae.py:331-366), ``ShopformerTransformer.forward``
(shopformer/models/transformer.py:304-329), ``Shopformer.compute_normality_score`` / ``forward``
(shopformer/models/shopformer.py:150-220), ``compute_anomaly_score`` (shopformer_2/models/shopformer.py:155-188),
``_normalize_sequence`` (shopformer/data/poselift_dataset.py:367-388).
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Tuple

import torch

from . import native as N
from .engine import ScoringEngine, _ptr, _stream_ptr

_engines: "weakref.WeakValueDictionary[int, ScoringEngine]" = weakref.WeakValueDictionary()


def model_handle(engine: ScoringEngine) -> int:
    """Integer handle of a packed model for the ``model`` argument of the ops (valid while the engine is alive)."""
    h = id(engine)
    _engines[h] = engine
    return h


def _engine(handle: int) -> ScoringEngine:
    try:
        return _engines[handle]
    except KeyError:
        raise RuntimeError(f"shopformer_b200: {handle} is not a live model handle (see ops.model_handle)") from None


_lib = torch.library.Library("shopformer_b200", "DEF")
_lib.define("window_normalize(Tensor raw, int num_keypoints, bool normalize) -> Tensor")
_lib.define("tokenize(Tensor poses, int model, str precision) -> Tensor")
_lib.define("reconstruct_tokens(Tensor tokens, int model, str precision) -> Tensor")
_lib.define("normality_score(Tensor tokens, Tensor recon, int model, str reduction) -> Tensor")
_lib.define("score_fused(Tensor poses, int model, str reduction, str precision) -> Tensor")
_lib.define("score_fused_full(Tensor poses, int model, str precision) -> (Tensor, Tensor, Tensor)")


def _window_normalize(raw: torch.Tensor, num_keypoints: int, normalize: bool) -> torch.Tensor:
    torch._check(raw.dim() == 4 and raw.shape[3] == 3, lambda: f"raw must be (B,T,K,3), got {tuple(raw.shape)}")
    x = raw.to(torch.float32).contiguous()
    B, T, K, _ = x.shape
    out = torch.empty(B, 2, T, num_keypoints, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.load().sf_normalize_windows(_ptr(x), B, T, K, num_keypoints, int(normalize), _ptr(out), _stream_ptr(x.device)),
                "sf_normalize_windows")
    return out


def _tokenize(poses: torch.Tensor, model: int, precision: str) -> torch.Tensor:
    return _engine(model).tokenize(poses, precision=precision)


def _reconstruct_tokens(tokens: torch.Tensor, model: int, precision: str) -> torch.Tensor:
    return _engine(model).reconstruct_tokens(tokens, precision=precision)


def _normality_score(tokens: torch.Tensor, recon: torch.Tensor, model: int, reduction: str) -> torch.Tensor:
    return _engine(model).normality_score(tokens, recon, reduction=reduction)


def _score_fused(poses: torch.Tensor, model: int, reduction: str, precision: str) -> torch.Tensor:
    return _engine(model).score_windows(poses, reduction=reduction, precision=precision)


def _score_fused_full(poses: torch.Tensor, model: int, precision: str):
    return _engine(model).score_windows(poses, precision=precision, return_tokens=True, return_recon=True)


_lib.impl("score_fused_full", _score_fused_full, "CUDA")
_lib.impl("window_normalize", _window_normalize, "CUDA")
_lib.impl("tokenize", _tokenize, "CUDA")
_lib.impl("reconstruct_tokens", _reconstruct_tokens, "CUDA")
_lib.impl("normality_score", _normality_score, "CUDA")
_lib.impl("score_fused", _score_fused, "CUDA")


# ---- shape functions (fake / meta tensors: torch.compile, opcheck)
def _token_shape(model: int, T: int) -> Tuple[int, int]:
    return _engine(model).token_shape(int(T))


@torch.library.register_fake("shopformer_b200::window_normalize")
def _(raw, num_keypoints, normalize):
    return raw.new_empty((raw.shape[0], 2, raw.shape[1], num_keypoints), dtype=torch.float32)


@torch.library.register_fake("shopformer_b200::tokenize")
def _(poses, model, precision):
    S, D = _token_shape(model, poses.shape[2])
    return poses.new_empty((poses.shape[0], S, D), dtype=torch.float32)


@torch.library.register_fake("shopformer_b200::reconstruct_tokens")
def _(tokens, model, precision):
    return tokens.new_empty(tokens.shape, dtype=torch.float32)


@torch.library.register_fake("shopformer_b200::normality_score")
def _(tokens, recon, model, reduction):
    shape = (tokens.shape[0],) if reduction == "mean" else (tokens.shape[0], tokens.shape[1])
    return tokens.new_empty(shape, dtype=torch.float32)


@torch.library.register_fake("shopformer_b200::score_fused")
def _(poses, model, reduction, precision):
    if reduction == "mean":
        return poses.new_empty((poses.shape[0],), dtype=torch.float32)
    S, _ = _token_shape(model, poses.shape[2])
    return poses.new_empty((poses.shape[0], S), dtype=torch.float32)


@torch.library.register_fake("shopformer_b200::score_fused_full")
def _(poses, model, precision):
    S, D = _token_shape(model, poses.shape[2])
    B = poses.shape[0]
    return (poses.new_empty((B,), dtype=torch.float32), poses.new_empty((B, S, D), dtype=torch.float32),
            poses.new_empty((B, S, D), dtype=torch.float32))


OPS = ("window_normalize", "tokenize", "reconstruct_tokens", "normality_score", "score_fused", "score_fused_full")
