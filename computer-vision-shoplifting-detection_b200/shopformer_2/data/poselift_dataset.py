"""PoseLift windows for the variant-2 drop-in (optional synthetic neck -> 18 keypoints,
per-window ``video_ids`` / ``frame_indices``, training-time augmentation).

Reference surface: shopformer_2/data/poselift_dataset.py (``add_neck_keypoint`` :57-91,
``PoseAugmentor`` :170-286, ``PoseLiftDataset`` :289-597, ``PoseLiftDataModule`` :600-676).

Window extraction (grouping aside) runs on the GPU through ``sf_window_normalize``; the neck
keypoint is synthesised inside that kernel.  The augmentor is training-only host code and is
kept API-compatible but outside the accelerated scope.
"""
import math
import os
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)

from shopformer_b200.engine import DeviceTracks, window_normalize  # noqa: E402
from shopformer_b200.ingest import load_poselift_split  # noqa: E402

__all__ = ["add_neck_keypoint", "PoseAugmentor", "PoseLiftDataset", "PoseLiftDataModule"]

_FLIP_PAIRS = [(1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12), (13, 14), (15, 16)]


def add_neck_keypoint(keypoints: np.ndarray) -> np.ndarray:
    """(>=17, C) -> (18, C): neck = shoulder midpoint, or the present shoulder, or zeros."""
    kp = keypoints
    if kp.shape[0] < 17:
        kp = np.vstack([kp, np.zeros((17 - kp.shape[0], kp.shape[1]))])
    ls, rs = kp[5], kp[6]
    missing_l, missing_r = np.allclose(ls[:2], 0), np.allclose(rs[:2], 0)
    if missing_l and missing_r:
        neck = np.zeros_like(ls)
    elif missing_l:
        neck = rs.copy()
    elif missing_r:
        neck = ls.copy()
    else:
        neck = (ls + rs) / 2.0
    return np.vstack([kp[:17], neck[None]])


class PoseAugmentor:
    """Random affine (flip / shear / scale / rotation / translation) + jitter + frame and keypoint
    dropout on a (T,V,C) window; draws from the global numpy RNG in the reference's order."""

    def __init__(self, flip_prob: float = 0.5, jitter_std: float = 0.02, scale_range: Tuple[float, float] = (0.9, 1.1),
                 rotation_range: float = 10.0, shear_range: float = 0.1, translation_range: float = 0.1,
                 temporal_dropout_prob: float = 0.1, keypoint_dropout_prob: float = 0.0, num_keypoints: int = 17):
        self.flip_prob, self.jitter_std, self.scale_range = flip_prob, jitter_std, scale_range
        self.rotation_range, self.shear_range, self.translation_range = rotation_range, shear_range, translation_range
        self.temporal_dropout_prob, self.keypoint_dropout_prob = temporal_dropout_prob, keypoint_dropout_prob
        self.num_keypoints = num_keypoints

    @classmethod
    def from_config(cls, config: Dict) -> "PoseAugmentor":
        a = config.get("data", {}).get("augmentation", {})
        return cls(flip_prob=a.get("flip_prob", 0.5), jitter_std=a.get("jitter_std", 0.02),
                   scale_range=tuple(a.get("scale_range", [0.9, 1.1])), rotation_range=a.get("rotation_range", 10.0),
                   shear_range=a.get("shear_range", 0.1), translation_range=a.get("translation_range", 0.1),
                   temporal_dropout_prob=a.get("temporal_dropout_prob", 0.1),
                   keypoint_dropout_prob=a.get("keypoint_dropout_prob", 0.0),
                   num_keypoints=config.get("model", {}).get("num_keypoints", 17))

    def __call__(self, pose_seq: np.ndarray) -> np.ndarray:
        out = pose_seq.copy()
        flip = np.random.random() < self.flip_prob
        shx = np.random.uniform(-self.shear_range, self.shear_range)
        shy = np.random.uniform(-self.shear_range, self.shear_range)
        sc = np.random.uniform(self.scale_range[0], self.scale_range[1])
        rot = math.radians(np.random.uniform(-self.rotation_range, self.rotation_range))
        tx = np.random.uniform(-self.translation_range, self.translation_range)
        ty = np.random.uniform(-self.translation_range, self.translation_range)
        c, s, f = math.cos(rot), math.sin(rot), (-1.0 if flip else 1.0)
        m = np.array([[sc * f * c - shy * sc * s, shx * sc * f * c - sc * s, tx * c - ty * s],
                      [sc * f * s + shy * sc * c, shx * sc * f * s + sc * c, tx * s + ty * c]], dtype=np.float32)
        xy1 = np.concatenate([pose_seq[:, :, :2], np.ones(pose_seq.shape[:2] + (1,))], axis=-1)
        out[:, :, :2] = xy1 @ m.T
        if flip:
            src = out.copy()
            for a, b in _FLIP_PAIRS:
                if a < self.num_keypoints and b < self.num_keypoints:
                    out[:, a], out[:, b] = src[:, b], src[:, a]
        if self.jitter_std > 0:
            out[:, :, :2] += np.random.randn(*out[:, :, :2].shape) * self.jitter_std
        if self.temporal_dropout_prob > 0:
            for t in range(out.shape[0]):
                if np.random.random() < self.temporal_dropout_prob:
                    out[t] = 0
        if self.keypoint_dropout_prob > 0:
            for t in range(out.shape[0]):
                for v in range(out.shape[1]):
                    if np.random.random() < self.keypoint_dropout_prob:
                        out[t, v] = 0
        return out


class PoseLiftDataset(Dataset):
    def __init__(self, data_dir: str, split: str = "train", seq_len: int = 24, stride: int = 12, num_keypoints: int = 17,
                 normalize: bool = True, include_confidence: bool = False, augmentor: Optional[PoseAugmentor] = None,
                 max_gap: int = 5, device: str = "cuda"):
        if include_confidence:
            raise NotImplementedError("include_confidence=True (3-channel windows) is not on the accelerated path")
        self.data_dir, self.split, self.seq_len, self.stride = data_dir, split, seq_len, stride
        self.num_keypoints, self.normalize, self.include_confidence = num_keypoints, normalize, False
        self.num_channels, self.augmentor, self.max_gap = 2, augmentor, max_gap
        tracks = load_poselift_split(data_dir, split)
        if not torch.cuda.is_available():
            raise RuntimeError("PoseLiftDataset windows poses on the GPU (sm_100a); no CUDA device is visible")
        out = window_normalize(DeviceTracks(tracks, torch.device(device)), seq_len, stride, num_keypoints=num_keypoints,
                               max_gap=max_gap, normalize=normalize, want_frame_indices=True)
        self.poses = out["poses"].cpu()                               # (N,2,T,V)
        self.labels: List[int] = out["labels"].cpu().tolist()
        track_of = out["window_track"].cpu().numpy()
        self.video_ids: List[str] = [tracks.video_names[tracks.track_video[t]] for t in track_of]
        self.frame_indices: List[List[int]] = out["frame_indices"].cpu().tolist()
        folder = "Train" if split == "train" else "Test"
        print(f"Loaded {len(self.labels)} sequences from {folder} split")
        if split == "test":
            pos = sum(self.labels)
            print(f"  Normal: {len(self.labels) - pos}, Anomaly: {pos}")

    @classmethod
    def from_config(cls, config: Dict, split: str = "train", augment: bool = True) -> "PoseLiftDataset":
        data, model = config.get("data", {}), config.get("model", {})
        aug = None
        if augment and split == "train" and data.get("augmentation", {}).get("enabled", True):
            aug = PoseAugmentor.from_config(config)
        return cls(data_dir=data.get("data_dir", "../shopformer/data/PoseLift"), split=split,
                   seq_len=model.get("seq_len", 24), stride=data.get("stride", 12),
                   num_keypoints=model.get("num_keypoints", 17), normalize=data.get("normalize", True),
                   include_confidence=data.get("include_confidence", False), augmentor=aug)

    @property
    def samples(self) -> List[np.ndarray]:
        return [p.permute(1, 2, 0).numpy() for p in self.poses]

    def __len__(self) -> int:
        return len(self.labels)

    def __getitem__(self, idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        pose = self.poses[idx]
        if self.augmentor is not None:
            tvc = self.augmentor(pose.permute(1, 2, 0).numpy().copy())
            pose = torch.from_numpy(np.ascontiguousarray(np.transpose(tvc, (2, 0, 1)))).float()
        else:
            pose = pose.clone()
        return pose, torch.tensor(self.labels[idx], dtype=torch.long)

    def get_video_info(self, idx: int) -> Dict:
        return {"video_id": self.video_ids[idx], "frame_indices": self.frame_indices[idx], "label": self.labels[idx]}


class PoseLiftDataModule:
    def __init__(self, config: Dict, num_workers: int = 0):
        self.config, self.num_workers = config, num_workers
        self.batch_size = config.get("training", {}).get("batch_size", 32)
        self.train_dataset: Optional[PoseLiftDataset] = None
        self.test_dataset: Optional[PoseLiftDataset] = None

    def setup(self):
        self.train_dataset = PoseLiftDataset.from_config(self.config, split="train", augment=True)
        self.test_dataset = PoseLiftDataset.from_config(self.config, split="test", augment=False)

    def train_dataloader(self) -> DataLoader:
        if self.train_dataset is None:
            raise RuntimeError("Call setup() before getting dataloaders")
        return DataLoader(self.train_dataset, batch_size=self.batch_size, shuffle=True, num_workers=self.num_workers,
                          pin_memory=False, drop_last=True)

    def test_dataloader(self) -> DataLoader:
        if self.test_dataset is None:
            raise RuntimeError("Call setup() before getting dataloaders")
        return DataLoader(self.test_dataset, batch_size=self.batch_size, shuffle=False, num_workers=self.num_workers,
                          pin_memory=False, drop_last=False)

    def get_stats(self) -> Dict[str, int]:
        stats: Dict[str, int] = {}
        if self.train_dataset:
            stats["train_samples"] = len(self.train_dataset)
        if self.test_dataset:
            pos = sum(self.test_dataset.labels)
            stats.update(test_samples=len(self.test_dataset), test_normal=len(self.test_dataset.labels) - pos, test_anomaly=pos)
        return stats
