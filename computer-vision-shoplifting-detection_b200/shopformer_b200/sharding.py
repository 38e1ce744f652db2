"""Window sharding across the GPUs of one box + the single collective of the path.

Windows are independent (eval-mode BatchNorm uses running statistics; no cross-window state),
so the path shards by contiguous window-index ranges with replicated weights (<= 10 MB) and no
data-path collective.  The only exchange is an all-gather of the fp32 scores over NVLink
(SURVEY 8e): every rank scores straight into its slice of the gather buffer, so no staging copy
precedes the collective.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous slice; slices are ceil(n/world) long, the last ones may be short/empty."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def padded_shard(n: int, world: int) -> int:
    return (n + world - 1) // world


class ShardedScorer:
    """Scores this rank's windows into its slice of a (world * per_rank,) buffer and all-gathers."""

    def __init__(self, engine, per_rank: int, group: Optional[dist.ProcessGroup] = None):
        self.engine = engine
        self.per_rank = per_rank
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        device = engine.device if engine is not None else torch.device("cpu")
        self.buffer = torch.zeros(self.world * per_rank, dtype=torch.float32, device=device)

    def my_slice(self) -> torch.Tensor:
        return self.buffer[self.rank * self.per_rank:(self.rank + 1) * self.per_rank]

    def gather(self) -> torch.Tensor:
        """In-place all-gather of every rank's slice (NCCL on GPUs, gloo in the CPU tests)."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.buffer, self.my_slice(), group=self.group)
        return self.buffer

    def score(self, poses: torch.Tensor, precision: str = "fp32") -> torch.Tensor:
        """poses: this rank's (<= per_rank, C, T, V) windows on the device -> all ranks' scores."""
        n = poses.shape[0]
        if n > self.per_rank:
            raise ValueError(f"{n} windows exceed the per-rank shard of {self.per_rank}")
        out = self.my_slice()
        if n < self.per_rank:
            out[n:].zero_()
        self.engine.score_windows(poses, precision=precision, out=out[:n])
        return self.gather()


def gather_scores(local: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Stand-alone helper: pad each rank's scores to ceil(n_total/world), all-gather, trim to n_total."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:n_total]
    per = padded_shard(n_total, world)
    send = torch.zeros(per, dtype=local.dtype, device=local.device)
    send[:local.numel()] = local
    out = torch.empty(world * per, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, send, group=group)
    return out[:n_total]


# --------------------------------------------------------------------------- sharding a sweep over packed tracks
def track_window_counts(track_offsets: np.ndarray, seq_len: int, stride: int) -> np.ndarray:
    """Candidate windows per track (positions 0, stride, ... while p + T <= len): the host-side count the reference's
    loop bound gives (shopformer/data/poselift_dataset.py:300-303)."""
    lens = np.diff(np.asarray(track_offsets, dtype=np.int64))
    return np.where(lens >= seq_len, (lens - seq_len) // stride + 1, 0).astype(np.int64)


def shard_tracks(track_offsets: np.ndarray, seq_len: int, stride: int, world: int) -> List[Tuple[int, int, int]]:
    """Contiguous track ranges of (nearly) equal candidate-window counts, one per rank: a host prefix sum of the per-track
    counts cut at multiples of total / world ("shard by video / track with a host prefix-sum so that the global
    window index = reference order").  Returns [(track_lo, track_hi, first_global_candidate)] per rank; windows never span
    tracks, so no halo is needed."""
    counts = track_window_counts(track_offsets, seq_len, stride)
    pre = np.concatenate([[0], np.cumsum(counts)])
    total = int(pre[-1])
    n_tracks = len(counts)
    cuts = [0]
    for r in range(1, world):
        target = (total * r + world - 1) // world
        t = int(np.searchsorted(pre, target, side="left"))
        cuts.append(min(max(t, cuts[-1]), n_tracks))
    cuts.append(n_tracks)
    return [(cuts[r], cuts[r + 1], int(pre[cuts[r]])) for r in range(world)]


def gather_ragged(local: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, List[int]]:
    """All-gather of per-rank score vectors of different lengths (a track shard's number of VALID windows is only known on
    its rank): lengths are exchanged first (one int64 per rank), scores are padded to the longest shard and gathered with
    one NCCL all-gather, then compacted in rank order = global window order.  Returns (all scores, per-rank counts)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local, [int(local.numel())]
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts_l = [int(c) for c in counts.tolist()]
    per = max(max(counts_l), 1)
    send = torch.zeros(per, dtype=local.dtype, device=local.device)
    send[:local.numel()] = local
    out = torch.empty(world * per, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, send, group=group)
    return torch.cat([out[r * per:r * per + counts_l[r]] for r in range(world)]), counts_l
