"""Canonical model configurations of the scoring path (SURVEY section 8 / BASELINE.md section 2).

  A   shopformer/   train.py defaults: V=17 T=24 H=32 d=136 S=3, 2 heads, 2+2 layers, ff 64
  A'  shopformer/   same with hidden_channels=64 (class default / sweep baseline)
  B   shopformer_2/ configs/paper_config.yaml: V=18 T=12 H=64 d=144 S=2, 2 heads, 2+2 layers, ff 64
  C   shopformer_2/ utils/config.get_default_config(): V=17 T=24 H=64, 136->144 projections,
      12 heads, 4+4 layers, ff 512
"""
from __future__ import annotations

import copy
from typing import Any, Dict

V1_CONFIGS: Dict[str, Dict[str, Any]] = {
    "A": dict(in_channels=2, hidden_channels=32, latent_channels=8, num_keypoints=17, seq_len=24, num_tokens=2,
              transformer_heads=2, transformer_layers=2, transformer_ff_dim=64, dropout=0.2),
    "A1": dict(in_channels=2, hidden_channels=64, latent_channels=8, num_keypoints=17, seq_len=24, num_tokens=2,
               transformer_heads=2, transformer_layers=2, transformer_ff_dim=64, dropout=0.2),
    # what evaluate.py/inference.py rebuild when config.json is absent (T=12 -> strides [2,2,1,1])
    "A12": dict(in_channels=2, hidden_channels=64, latent_channels=8, num_keypoints=17, seq_len=12, num_tokens=2,
                transformer_heads=2, transformer_layers=2, transformer_ff_dim=64, dropout=0.1),
}

_V2_MODEL_B = {
    "in_channels": 2, "num_keypoints": 18, "seq_len": 12, "num_tokens": 2,
    "gcae": {"hidden_channels": 64, "latent_channels": 8, "num_layers": 4, "dropout": 0.1},
    "transformer": {"input_dim": 144, "d_model": 144, "num_heads": 2, "num_layers": 2, "dim_feedforward": 64,
                    "dropout": 0.1},
}
_V2_MODEL_C = {
    "in_channels": 2, "num_keypoints": 17, "seq_len": 24, "num_tokens": 2,
    "gcae": {"hidden_channels": 64, "latent_channels": 8, "num_layers": 4, "dropout": 0.1},
    "transformer": {"input_dim": 136, "d_model": 144, "num_heads": 12, "num_layers": 4, "dim_feedforward": 512,
                    "dropout": 0.1},
}
# exercises the AdaptiveAvgPool branch: 24 // 5 = 4 -> strides [2,2,1,1], 24->12->6->6->6, pooled 6 -> 5
_V2_MODEL_P = {
    "in_channels": 2, "num_keypoints": 17, "seq_len": 24, "num_tokens": 5,
    "gcae": {"hidden_channels": 32, "latent_channels": 8, "num_layers": 4, "dropout": 0.1},
    "transformer": {"input_dim": 136, "d_model": 136, "num_heads": 4, "num_layers": 2, "dim_feedforward": 128,
                    "dropout": 0.1},
}
V2_CONFIGS: Dict[str, Dict[str, Any]] = {"B": {"model": _V2_MODEL_B}, "C": {"model": _V2_MODEL_C},
                                         "P": {"model": _V2_MODEL_P}}

ALL_CONFIGS = ("A", "A1", "A12", "B", "C", "P")


def variant_of(name: str) -> int:
    return 1 if name in V1_CONFIGS else 2


def input_shape(name: str):
    """(C, T, V) of one window."""
    if name in V1_CONFIGS:
        c = V1_CONFIGS[name]
        return c["in_channels"], c["seq_len"], c["num_keypoints"]
    m = V2_CONFIGS[name]["model"]
    return m["in_channels"], m["seq_len"], m["num_keypoints"]


def ctor_args(name: str):
    return copy.deepcopy(V1_CONFIGS[name] if name in V1_CONFIGS else V2_CONFIGS[name])


# useful FLOPs per window on the scoring path (excludes zero-padding taps and
# the GCAE pose decoder) -- the numerator of the tensor roofline
USEFUL_FLOPS = {"A": 10_263_648, "A1": 30_209_952, "B": 10_028_736, "C": 25_941_344}
DENSE_FLOPS = {"A": 11_347_296, "A1": 34_427_040, "B": 14_632_128, "C": 29_731_936}
