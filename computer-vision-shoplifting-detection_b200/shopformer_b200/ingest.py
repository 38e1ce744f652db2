"""One-time host conversion: PoseLift pickles -> packed per-person tracks.

The GPU windowing kernel needs, per person, the detections sorted by frame number in one
contiguous array, and the persons in the order the reference would visit them.  That order
is data dependent -- first appearance while iterating the pickled frame dict
(shopformer/data/poselift_dataset.py:267-298; shopformer_2/...:425-453) -- so it is fixed
here, on the host, exactly once; everything downstream (window enumeration, continuity,
labels, gather, normalisation) runs on the device.
"""
from __future__ import annotations

import pickle
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .engine import PackedTracks


def _detection_rows(kp: np.ndarray, k_out: int) -> np.ndarray:
    """(k_out, 3) fp32 rows of one detection: flat arrays are (-1,3)-reshaped, short ones
    zero-padded, (n,2) arrays get a zero confidence column."""
    if kp.ndim == 1:
        kp = kp.reshape(-1, 3)
    rows = np.zeros((k_out, 3), dtype=np.float32)
    n, c = min(kp.shape[0], k_out), min(kp.shape[1], 3)
    rows[:n, :c] = kp[:n, :c]
    return rows


def pack_videos(videos: Sequence[Tuple[str, Dict[Any, Any], Optional[np.ndarray]]], kp_per_frame: int = 17) -> PackedTracks:
    """videos: (name, pose_dict, frame_labels or None) in the order the reference loads them
    (sorted pickle file names).  Skips exactly what the reference skips: empty / non-dict
    frames, malformed person records, detections containing NaN or Inf."""
    kps: List[np.ndarray] = []
    fnos: List[np.ndarray] = []
    offs = [0]
    tvid: List[int] = []
    gts: List[np.ndarray] = []
    gt_offs = [0]
    names: List[str] = []
    any_gt = False
    for vid, (name, pose_data, labels) in enumerate(videos):
        names.append(name)
        persons: Dict[Any, Dict[int, np.ndarray]] = {}
        for frame_num, frame_data in pose_data.items():
            if not frame_data or not isinstance(frame_data, dict):
                continue
            for pid, rec in frame_data.items():
                if not isinstance(rec, (list, tuple)) or len(rec) < 2:
                    continue
                kp = np.array(rec[1])
                if np.any(np.isnan(kp)) or np.any(np.isinf(kp)):
                    continue
                persons.setdefault(pid, {})[int(frame_num)] = kp
        for _pid, frames in persons.items():
            order = sorted(frames.keys())
            kps.append(np.stack([_detection_rows(frames[f], kp_per_frame) for f in order]) if order
                       else np.zeros((0, kp_per_frame, 3), np.float32))
            fnos.append(np.asarray(order, dtype=np.int32))
            offs.append(offs[-1] + len(order))
            tvid.append(vid)
        g = np.zeros(0, np.uint8) if labels is None else np.asarray(labels).astype(np.uint8).reshape(-1)
        any_gt |= labels is not None
        gts.append(g)
        gt_offs.append(gt_offs[-1] + len(g))
    kp_all = np.concatenate(kps, axis=0) if kps else np.zeros((0, kp_per_frame, 3), np.float32)
    fn_all = np.concatenate(fnos) if fnos else np.zeros(0, np.int32)
    return PackedTracks(kp=np.ascontiguousarray(kp_all, dtype=np.float32), frame_no=fn_all,
                        track_offsets=np.asarray(offs, dtype=np.int64), track_video=np.asarray(tvid, dtype=np.int32),
                        gt=np.concatenate(gts) if any_gt else None,
                        gt_offsets=np.asarray(gt_offs, dtype=np.int64) if any_gt else None, video_names=names)


def load_poselift_split(data_dir: str, split: str = "train", kp_per_frame: int = 17) -> PackedTracks:
    """Pickle_files/{Train|Test}/*.pkl (+ Pickle_files/GT/<video>.npy for the test split),
    file order and GT lookup as shopformer/data/poselift_dataset.py:231-254."""
    root = Path(data_dir)
    pose_dir = root / "Pickle_files" / ("Train" if split == "train" else "Test")
    if not pose_dir.exists():
        raise FileNotFoundError(f"Pose directory not found: {pose_dir}")
    label_dir = root / "Pickle_files" / "GT" if split == "test" else None
    videos = []
    for pkl in sorted(pose_dir.glob("*.pkl")):
        with open(pkl, "rb") as f:
            data = pickle.load(f)
        labels = None
        if label_dir is not None and (label_dir / f"{pkl.stem}.npy").exists():
            labels = np.load(label_dir / f"{pkl.stem}.npy")
        videos.append((pkl.stem, data, labels))
    return pack_videos(videos, kp_per_frame=kp_per_frame)
