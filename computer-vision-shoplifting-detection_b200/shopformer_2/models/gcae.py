"""Variant-2 GCAE tokenizer shells over ``shopformer_b200.modules``.

Reference surface: shopformer_2/models/gcae.py (adjacency :22-99, blocks :102-270, ``GCAEEncoder``
:273-422, ``GCAEDecoder`` :425-534, ``GCAE`` :537-613).  Graph: COCO-17, or COCO+neck whenever
``num_keypoints == 18``; strides from the factorisation of ``seq_len // num_tokens``; adaptive
average pooling when the strided lengths do not land on ``num_tokens``.
"""
from typing import Tuple

import numpy as np

from shopformer_b200 import modules as _m
from shopformer_b200.modules import GraphConvolution, STGCNBlock, TemporalConvolution, normalize_adjacency

__all__ = ["get_skeleton_adjacency", "normalize_adjacency", "GraphConvolution", "TemporalConvolution",
           "STGCNBlock", "GCAEEncoder", "GCAEDecoder", "GCAE"]

_FAMILY = 2


def get_skeleton_adjacency(num_keypoints: int = 17, layout: str = "coco") -> np.ndarray:
    return _m.get_skeleton_adjacency(num_keypoints, layout, _FAMILY)


class GCAEEncoder(_m.GCAEEncoder):
    def __init__(self, in_channels: int = 2, hidden_channels: int = 64, out_channels: int = 8,
                 num_keypoints: int = 17, seq_len: int = 24, num_tokens: int = 2, num_layers: int = 4,
                 dropout: float = 0.1, layout: str = "coco"):
        super().__init__(in_channels, hidden_channels, out_channels, num_keypoints, seq_len, num_tokens,
                         num_layers, dropout, layout, _FAMILY)

    def _compute_strides(self, seq_len: int, num_tokens: int, num_layers: int) -> list:
        strides, self._needs_pooling = _m.strides_factorised(seq_len, num_tokens, num_layers)
        return strides


class GCAEDecoder(_m.GCAEDecoder):
    def __init__(self, in_channels: int = 8, hidden_channels: int = 64, out_channels: int = 2,
                 num_keypoints: int = 17, seq_len: int = 24, num_tokens: int = 2, num_layers: int = 4,
                 dropout: float = 0.1, layout: str = "coco"):
        super().__init__(in_channels, hidden_channels, out_channels, num_keypoints, seq_len, num_tokens,
                         num_layers, dropout, layout, _FAMILY)

    def _compute_upsample_factors(self, num_tokens: int, seq_len: int, num_layers: int) -> list:
        return _m.upsample_factors(num_tokens, seq_len, num_layers)


class GCAE(_m.GCAE):
    def __init__(self, in_channels: int = 2, hidden_channels: int = 64, latent_channels: int = 8,
                 num_keypoints: int = 17, seq_len: int = 24, num_tokens: int = 2, num_layers: int = 4,
                 dropout: float = 0.1, layout: str = "coco"):
        super().__init__(in_channels, hidden_channels, latent_channels, num_keypoints, seq_len, num_tokens,
                         num_layers, dropout, layout, _FAMILY)
