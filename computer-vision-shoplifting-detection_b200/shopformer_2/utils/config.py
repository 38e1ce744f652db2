"""YAML config helpers, format-compatible with shopformer_2/utils/config.py:12-202 (host side,
outside the accelerated scope; ``get_default_config()`` defines canonical config C)."""
import copy
import os
from pathlib import Path
from typing import Any, Dict

import yaml

_DEFAULT = {
    "model": {
        "in_channels": 2, "num_keypoints": 17, "seq_len": 24, "num_tokens": 2,
        "gcae": {"hidden_channels": 64, "latent_channels": 8, "num_layers": 4, "dropout": 0.1},
        "transformer": {"input_dim": 136, "d_model": 144, "num_heads": 12, "num_layers": 4,
                        "dim_feedforward": 512, "dropout": 0.1},
    },
    "training": {
        "device": "auto",
        "stage1": {"epochs": 50, "learning_rate": 5e-5, "weight_decay": 1e-4},
        "stage2": {"epochs": 100, "learning_rate": 5e-5, "weight_decay": 1e-4},
        "batch_size": 32, "gradient_accumulation": 4, "grad_clip": 1.0,
        "scheduler": {"type": "cosine_warmup", "warmup_epochs": 5, "min_lr": 1e-6},
        "early_stopping": {"enabled": True, "patience": 20, "min_delta": 0.001},
    },
    "data": {
        "data_dir": "../shopformer/data/PoseLift", "stride": 12, "normalize": True,
        "augmentation": {"enabled": True, "flip_prob": 0.5, "jitter_std": 0.02, "scale_range": [0.9, 1.1],
                         "rotation_range": 10.0},
    },
}


def load_config(config_path: str) -> Dict[str, Any]:
    path = Path(config_path)
    if not path.exists():
        raise FileNotFoundError(f"Config file not found: {path}")
    with open(path, "r") as f:
        cfg = yaml.safe_load(f)
    data = cfg.get("data") if isinstance(cfg, dict) else None
    if data and "data_dir" in data and not os.path.isabs(data["data_dir"]):
        data["data_dir"] = str((path.parent / data["data_dir"]).resolve())
    return cfg


def save_config(config: Dict[str, Any], save_path: str):
    path = Path(save_path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as f:
        yaml.dump(config, f, default_flow_style=False, sort_keys=False)


def merge_configs(base_config: Dict, override_config: Dict) -> Dict:
    out = base_config.copy()
    for k, v in override_config.items():
        out[k] = merge_configs(out[k], v) if isinstance(out.get(k), dict) and isinstance(v, dict) else v
    return out


def get_default_config() -> Dict[str, Any]:
    return copy.deepcopy(_DEFAULT)


def validate_config(config: Dict[str, Any]) -> bool:
    need = {"model": ["in_channels", "num_keypoints", "seq_len", "num_tokens", "gcae", "transformer"],
            "training": ["stage1", "stage2", "batch_size"], "data": ["data_dir"]}
    for sec, keys in need.items():
        if sec not in config:
            raise ValueError(f"Missing required section: {sec}")
        for k in keys:
            if k not in config[sec]:
                raise ValueError(f"Missing required field: {sec}.{k}")
    return True
