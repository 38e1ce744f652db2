"""Deterministic synthetic weights, windows and PoseLift-style tracks.

Nothing here depends on torch's RNG stream or on module construction order: every
state-dict tensor is filled from ``np.random.RandomState(crc32(key) ^ seed)`` (the
frozen legacy generator), so the reference model in the build container and this
package on the GPU box get bit-identical weights from nothing but the key names.

Window statistics follow the reference's ``SyntheticPoseLiftDataset``
(shopformer/data/poselift_dataset.py:431-453): a COCO-17 template skeleton in [0,1]^2,
N(0,0.02) shape noise, per-frame N(0,0.02) motion (normal) or N(0,0.08) plus a
wrist->hip blend in the second half (anomalous).  The reference generator is a
per-sample Python loop on the unseeded global numpy RNG; this one is vectorised and
seeded (SURVEY 8d allows exactly that for N > 65 536).
"""
from __future__ import annotations

import zlib
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

COCO17_TEMPLATE = np.array([
    [0.5, 0.1], [0.48, 0.08], [0.52, 0.08], [0.45, 0.1], [0.55, 0.1],
    [0.4, 0.25], [0.6, 0.25], [0.35, 0.4], [0.65, 0.4], [0.3, 0.55],
    [0.7, 0.55], [0.45, 0.55], [0.55, 0.55], [0.43, 0.75], [0.57, 0.75],
    [0.42, 0.95], [0.58, 0.95]], dtype=np.float64)


# --------------------------------------------------------------------------- weights
def _rs(key: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0xFFFFFFFF)


def synth_state_dict(template: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Return a state dict with the template's keys/shapes/dtypes and deterministic
    values.  Buffers that are pure functions of the config (``*.adj``, ``*.pe``) and the
    integer ``num_batches_tracked`` counters are kept as they are."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in template.items():
        leaf = k.rsplit(".", 1)[-1]
        if leaf in ("adj", "pe", "num_batches_tracked") or not v.is_floating_point():
            out[k] = v.detach().clone()
            continue
        r = _rs(k, seed)
        shape = tuple(v.shape)
        if leaf == "running_var":
            a = r.uniform(0.5, 1.5, shape)
        elif leaf == "running_mean":
            a = r.normal(0.0, 0.1, shape)
        elif v.dim() == 1 and leaf == "weight":          # BatchNorm / LayerNorm scale
            a = r.uniform(0.8, 1.2, shape)
        elif v.dim() == 1:                               # biases (incl. in_proj_bias)
            a = r.normal(0.0, 0.05, shape)
        else:                                            # xavier-uniform-like
            rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
            bound = np.sqrt(6.0 / ((shape[0] + shape[1]) * rf))
            a = r.uniform(-bound, bound, shape)
        out[k] = torch.from_numpy(np.ascontiguousarray(a)).to(v.dtype)
    return out


# --------------------------------------------------------------------------- windows
def synth_windows(n: int, seq_len: int = 24, num_keypoints: int = 17, anomaly_ratio: float = 0.3,
                  seed: int = 1234) -> Tuple[np.ndarray, np.ndarray]:
    """(n, 2, T, V) fp32 windows + int64 labels (1 = anomalous)."""
    r = np.random.RandomState(seed)
    lab = (r.random_sample(n) < anomaly_ratio)
    base = COCO17_TEMPLATE[None] + r.randn(n, 17, 2) * 0.02
    scale = np.where(lab, 0.08, 0.02)[:, None, None, None]
    seq = base[:, None] + r.randn(n, seq_len, 17, 2) * scale
    late = (np.arange(seq_len) > seq_len // 2)[None, :, None] & lab[:, None, None]
    for w, h in ((9, 11), (10, 12)):
        blend = seq[:, :, w] * 0.7 + seq[:, :, h] * 0.3
        seq[:, :, w] = np.where(late, blend, seq[:, :, w])
    if num_keypoints == 18:                                # synthetic neck = shoulder midpoint
        neck = 0.5 * (seq[:, :, 5] + seq[:, :, 6])
        seq = np.concatenate([seq, neck[:, :, None]], axis=2)
    elif num_keypoints != 17:
        seq = seq[:, :, :num_keypoints]
    x = np.ascontiguousarray(np.transpose(seq, (0, 3, 1, 2)), dtype=np.float32)
    return x, lab.astype(np.int64)


# --------------------------------------------------------------------------- PoseLift-style fixtures
def synth_poselift_video(seed: int, n_frames: int = 300, *, dtype=np.float64) -> Tuple[Dict[Any, Any], np.ndarray]:
    """One fabricated PoseLift video dict ``{frame: {person: [bbox, kp(17,3)]}}`` plus its
    per-frame GT array, exercising every branch of the reference's ingest
    (SURVEY 8c-1): a gap > 5 frames, a short (< T) track, persons first seen in reverse
    id order, an empty frame, NaN / Inf keypoints, all-zero keypoints, a flat (51,)
    array, a 15-keypoint array, string frame keys and non-monotone insertion order."""
    r = np.random.RandomState(seed)
    frames: Dict[Any, Any] = {}

    def person_kp(f: int, pid: int) -> np.ndarray:
        base = COCO17_TEMPLATE * np.array([400.0, 800.0]) + np.array([300.0 + 40 * pid + 1.5 * f, 100.0])
        xy = base + r.randn(17, 2) * 4.0
        conf = r.uniform(0.2, 1.0, (17, 1))
        return np.concatenate([xy, conf], axis=1).astype(dtype)

    order = list(range(n_frames))
    order[10], order[11] = order[11], order[10]           # non-monotone dict insertion order
    for f in order:
        entry: Dict[Any, Any] = {}
        if 50 <= f < 130:                                  # person 7 first seen before person 1? no: 1 starts at 0
            entry[7] = [np.array([0, 0, 10, 10.0]), person_kp(f, 7)]
        if not (100 <= f < 108):                           # person 1: 8-frame gap
            entry[1] = [np.array([0, 0, 10, 10.0]), person_kp(f, 1)]
        if 200 <= f < 215:                                 # person 3: 15 frames only
            entry[3] = [None, person_kp(f, 3)]
        if f == 20:
            entry[1][1][4, :2] = 0.0                       # one zero (invalid) keypoint
        if f == 21:
            entry[1][1] = entry[1][1].reshape(-1)          # flat (51,) layout
        if f == 22:
            entry[1][1] = entry[1][1][:15]                 # fewer than 17 keypoints
        if f == 60:
            entry[7][1][3, 0] = np.nan                     # dropped detection (NaN)
        if f == 61:
            entry[7][1][2, 1] = np.inf                     # dropped detection (Inf)
        if f == 150:
            entry = {}                                     # empty frame
        if f == 151:
            entry[1] = [np.zeros(4)]                       # malformed person record
        frames[str(f) if f % 2 else f] = entry             # mix of str and int keys
    # a second video-late person with a *smaller* id first seen after a larger one
    for f in range(240, 290):
        key = str(f) if f % 2 else f
        frames[key][0] = [np.zeros(4), person_kp(f, 0) * (0.0 if f == 250 else 1.0)]  # f=250 all-zero pose
    gt = np.zeros(n_frames - 20, dtype=np.int64)           # shorter than the video: exercises min(f, len-1)
    gt[120:200] = 1
    gt[-1] = 1
    return frames, gt


def synth_tracks(n_tracks: int, seed: int = 1234, *, min_len: int = 30, max_len: int = 3000,
                 drop_prob: float = 0.02, zero_kp_prob: float = 0.01, gap_every: int = 500
                 ) -> Dict[str, np.ndarray]:
    """Packed tracks for the windowing kernel benchmark: AoS keypoints
    ``kp[F,17,3]`` fp32 in pixel units, ``frame_no[F]`` int32, ``track_offsets[n+1]`` int64,
    one GT array per track (each track is its own video here)."""
    r = np.random.RandomState(seed)
    lens = r.randint(min_len, max_len + 1, n_tracks)
    kps: List[np.ndarray] = []
    fnos: List[np.ndarray] = []
    gts: List[np.ndarray] = []
    offs = [0]
    gt_offs = [0]
    for n in lens:
        fn = np.arange(n, dtype=np.int64)
        keep = r.random_sample(n) >= drop_prob
        keep[0] = True
        fn = fn[keep]
        if gap_every > 0:                                  # a 6..20-frame hole every ~gap_every frames
            n_gaps = len(fn) // gap_every
            for g in range(n_gaps):
                at = (g + 1) * gap_every - r.randint(0, gap_every // 4 + 1)
                fn[at:] += r.randint(6, 21)
        m = len(fn)
        base = COCO17_TEMPLATE * np.array([300.0, 700.0]) + r.uniform(0, [1500.0, 300.0], (1, 2))
        xy = base[None] + np.cumsum(r.randn(m, 1, 2) * 1.5, axis=0) + r.randn(m, 17, 2) * 3.0
        conf = r.uniform(0.1, 1.0, (m, 17, 1))
        kp = np.concatenate([xy, conf], axis=2)
        zero = r.random_sample((m, 17)) < zero_kp_prob
        kp[zero] = 0.0
        kps.append(kp.astype(np.float32))
        fnos.append(fn.astype(np.int32))
        gt = np.zeros(int(fn[-1]) + 1, dtype=np.uint8)
        a = r.randint(0, len(gt))
        gt[a:a + r.randint(10, 200)] = 1
        gts.append(gt)
        offs.append(offs[-1] + m)
        gt_offs.append(gt_offs[-1] + len(gt))
    return {
        "kp": np.ascontiguousarray(np.concatenate(kps, axis=0)),
        "frame_no": np.concatenate(fnos),
        "track_offsets": np.asarray(offs, dtype=np.int64),
        "track_video": np.arange(n_tracks, dtype=np.int32),
        "gt": np.concatenate(gts),
        "gt_offsets": np.asarray(gt_offs, dtype=np.int64),
    }


def synth_tracks_device(n_tracks: int, track_len: int, device, seed: int = 1234, *, gap_every: int = 500, channels: int = 3,
                        chunk_tracks: int = 2048):
    """Bench-scale packed tracks generated ON the device (a 10 M-window sweep is ~120 M detections = 24.5 GB of
    keypoints; numpy would take minutes): `n_tracks` tracks of `track_len` detections each, COCO-17 skeleton doing a random
    walk in pixel units, a hole of 6..20 frames every ~`gap_every` detections (windows across it fail the continuity
    test), one ground-truth array per track.  Returns (kp (F,17,channels) fp32 cuda, frame_no (F,) int32 cuda,
    track_offsets (n+1,) int64 numpy, track_video (n,) int32 numpy, gt (sum,) uint8 cuda, gt_offsets (n+1,) int64 numpy)."""
    import torch
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    F = n_tracks * track_len
    kp = torch.empty(F, 17, channels, dtype=torch.float32, device=dev)
    frame_no = torch.empty(F, dtype=torch.int32, device=dev)
    tmpl = torch.from_numpy((COCO17_TEMPLATE * np.array([300.0, 700.0])).astype(np.float32)).to(dev)
    n_gaps = track_len // gap_every if gap_every > 0 else 0
    gt_len = track_len + 21 * max(n_gaps, 1)
    gt = torch.zeros(n_tracks * gt_len, dtype=torch.uint8, device=dev)
    for t0 in range(0, n_tracks, chunk_tracks):
        n = min(chunk_tracks, n_tracks - t0)
        base = torch.rand(n, 1, 1, 2, device=dev, generator=g) * torch.tensor([1500.0, 300.0], device=dev)
        walk = torch.cumsum(torch.randn(n, track_len, 1, 2, device=dev, generator=g) * 1.5, dim=1)
        xy = tmpl.view(1, 1, 17, 2) + base + walk + torch.randn(n, track_len, 17, 2, device=dev, generator=g) * 3.0
        view = kp[t0 * track_len:(t0 + n) * track_len].view(n, track_len, 17, channels)
        view[..., :2] = xy
        if channels == 3:
            view[..., 2] = torch.rand(n, track_len, 17, device=dev, generator=g) * 0.9 + 0.1
        fn = torch.arange(track_len, device=dev, dtype=torch.int32).repeat(n, 1)
        for k in range(n_gaps):
            at = (k + 1) * gap_every - torch.randint(0, gap_every // 4 + 1, (n, 1), device=dev, generator=g)
            hole = torch.randint(6, 21, (n, 1), device=dev, generator=g, dtype=torch.int32)
            fn += (torch.arange(track_len, device=dev).view(1, -1) >= at).to(torch.int32) * hole
        frame_no[t0 * track_len:(t0 + n) * track_len] = fn.reshape(-1)
        a = torch.randint(0, track_len, (n, 1), device=dev, generator=g)
        ln = torch.randint(10, 200, (n, 1), device=dev, generator=g)
        pos = torch.arange(gt_len, device=dev).view(1, -1)
        gt[t0 * gt_len:(t0 + n) * gt_len] = ((pos >= a) & (pos < a + ln)).to(torch.uint8).reshape(-1)
    track_offsets = np.arange(n_tracks + 1, dtype=np.int64) * track_len
    gt_offsets = np.arange(n_tracks + 1, dtype=np.int64) * gt_len
    return kp, frame_no, track_offsets, np.arange(n_tracks, dtype=np.int32), gt, gt_offsets
