"""PoseLift windows for the variant-1 drop-in.

Reference surface: shopformer/data/poselift_dataset.py (``PoseLiftDataset`` :200-400,
``SyntheticPoseLiftDataset`` :403-464, ``PoseLiftDataModule`` :467-533).

``PoseLiftDataset.__init__`` does what the reference does in nested Python loops -- group
detections per person, slide windows, test continuity, vote the label, gather and
centre/scale-normalise -- but on the GPU: one host pass packs the pickles into per-person
tracks (`shopformer_b200.ingest`), then ``sf_window_normalize`` produces every window in one
launch sequence.  ``samples`` / ``labels`` are exposed like the reference's ((T,V,2) fp32
arrays and ints) and ``__getitem__`` returns the same ``(2,T,V)`` float tensor + long label.
"""
import os
import sys
from typing import Tuple

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)

from shopformer_b200.engine import DeviceTracks, window_normalize  # noqa: E402
from shopformer_b200.ingest import load_poselift_split  # noqa: E402
from shopformer_b200.synthetic import COCO17_TEMPLATE  # noqa: E402

__all__ = ["PoseLiftDataset", "SyntheticPoseLiftDataset", "PoseLiftDataModule"]


class _WindowStore(Dataset):
    """Windows kept as one (N,C,T,V) CPU tensor; items are views of it."""
    poses: torch.Tensor
    labels: list

    def __len__(self) -> int:
        return len(self.labels)

    def __getitem__(self, idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.poses[idx].clone(), torch.tensor(self.labels[idx], dtype=torch.long)

    @property
    def samples(self):
        """Reference-format view: list of (T,V,C) fp32 arrays."""
        return [p.permute(1, 2, 0).numpy() for p in self.poses]


class PoseLiftDataset(_WindowStore):
    def __init__(self, data_dir: str, split: str = "train", seq_len: int = 12, stride: int = 6,
                 num_keypoints: int = 17, normalize: bool = True, include_confidence: bool = False,
                 device: str = "cuda"):
        self.data_dir, self.split, self.seq_len, self.stride = data_dir, split, seq_len, stride
        self.num_keypoints, self.normalize, self.include_confidence = num_keypoints, normalize, bool(include_confidence)
        self.num_channels = 3 if include_confidence else 2
        # variant 1 takes the detection's own first `num_keypoints` rows and zero-pads (poselift_dataset.py:345-354):
        # keep that many source rows per detection; it never synthesises a neck
        tracks = load_poselift_split(data_dir, split, kp_per_frame=max(17, num_keypoints))   # raises FileNotFoundError like the reference
        if not torch.cuda.is_available():
            raise RuntimeError("PoseLiftDataset windows poses on the GPU (sm_100a); no CUDA device is visible")
        dev = DeviceTracks(tracks, torch.device(device))
        out = window_normalize(dev, seq_len, stride, num_keypoints=num_keypoints, max_gap=5, normalize=normalize,
                               add_neck=False, include_confidence=self.include_confidence)
        self.poses = out["poses"].cpu()
        self.labels = out["labels"].cpu().tolist()
        self.window_track = out["window_track"].cpu()
        self.window_start = out["window_start"].cpu()


class SyntheticPoseLiftDataset(_WindowStore):
    """Random skeleton windows for smoke runs.  Draws from the global numpy RNG in the same
    order as the reference generator, so a caller who seeds numpy gets the same windows."""

    def __init__(self, num_samples: int = 1000, seq_len: int = 12, num_keypoints: int = 17,
                 num_channels: int = 2, anomaly_ratio: float = 0.3):
        self.num_samples, self.seq_len = num_samples, seq_len
        self.num_keypoints, self.num_channels = num_keypoints, num_channels
        wins = np.empty((num_samples, seq_len, num_keypoints, 2), dtype=np.float64)
        self.labels = []
        for i in range(num_samples):
            anomalous = np.random.random() < anomaly_ratio
            base = (COCO17_TEMPLATE + np.random.randn(17, 2) * 0.02)[:num_keypoints]
            sigma = 0.08 if anomalous else 0.02
            for t in range(seq_len):
                pose = base + np.random.randn(*base.shape) * sigma
                if anomalous and t > seq_len // 2:
                    pose[9] = pose[9] * 0.7 + pose[11] * 0.3
                    pose[10] = pose[10] * 0.7 + pose[12] * 0.3
                wins[i, t] = pose
            self.labels.append(1 if anomalous else 0)
        self.poses = torch.from_numpy(np.ascontiguousarray(np.transpose(wins, (0, 3, 1, 2)))).float()

    @property
    def samples(self):
        return [p.permute(1, 2, 0).double().numpy() for p in self.poses]


class PoseLiftDataModule:
    def __init__(self, data_dir: str, batch_size: int = 32, seq_len: int = 12, stride: int = 6,
                 num_workers: int = 4, use_synthetic: bool = False, synthetic_samples: int = 1000):
        self.data_dir, self.batch_size, self.seq_len, self.stride = data_dir, batch_size, seq_len, stride
        self.num_workers, self.use_synthetic, self.synthetic_samples = num_workers, use_synthetic, synthetic_samples
        self.train_dataset = None
        self.test_dataset = None

    def setup(self):
        if self.use_synthetic:
            self.train_dataset = SyntheticPoseLiftDataset(self.synthetic_samples, self.seq_len, anomaly_ratio=0.0)
            self.test_dataset = SyntheticPoseLiftDataset(self.synthetic_samples // 5, self.seq_len, anomaly_ratio=0.3)
        else:
            self.train_dataset = PoseLiftDataset(self.data_dir, "train", self.seq_len, self.stride)
            self.test_dataset = PoseLiftDataset(self.data_dir, "test", self.seq_len, self.stride)

    def _loader(self, ds, shuffle: bool) -> DataLoader:
        return DataLoader(ds, batch_size=self.batch_size, shuffle=shuffle, num_workers=self.num_workers, pin_memory=True)

    def train_dataloader(self) -> DataLoader:
        return self._loader(self.train_dataset, True)

    def test_dataloader(self) -> DataLoader:
        return self._loader(self.test_dataset, False)
