"""CPU, world_size 2 over gloo: the N>1 host logic (shard ranges, padding, in-place score all-gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shopformer_b200.sharding import ShardedScorer, gather_scores, padded_shard, shard_range


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 8, 10_000_001):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) <= padded_shard(n, world)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeEngine:
    """Stands in for ScoringEngine on CPU: score = mean of the window (deterministic, shard-independent)."""
    device = torch.device("cpu")

    def score_windows(self, poses, precision="fp32", out=None):
        out.copy_(poses.mean(dim=(1, 2, 3)))
        return out


def _worker(rank: int, world: int, port: int, n_total: int):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        poses = torch.randn(n_total, 2, 4, 3, generator=g)
        want = poses.mean(dim=(1, 2, 3))
        lo, hi = shard_range(n_total, rank, world)
        per = padded_shard(n_total, world)
        scorer = ShardedScorer(_FakeEngine(), per)
        allv = scorer.score(poses[lo:hi])
        assert allv.shape == (world * per,)
        assert torch.equal(allv[:n_total], want)              # global window order == reference order
        assert torch.equal(gather_scores(want[lo:hi].clone(), n_total), want)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 11])
def test_sharded_score_allgather_world2(n_total):
    mp.spawn(_worker, args=(2, _free_port(), n_total), nprocs=2, join=True)


# ---- sharding a sweep over packed tracks (BASELINE configs[3]: windows shard by video / track and window index)
def test_shard_tracks_balances_and_covers():
    import numpy as np
    from shopformer_b200.sharding import shard_tracks, track_window_counts
    rs = np.random.RandomState(0)
    lens = rs.randint(1, 3000, 500)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    counts = track_window_counts(off, 24, 12)
    assert counts[lens < 24].sum() == 0 and counts.sum() == sum((n - 24) // 12 + 1 for n in lens if n >= 24)
    for world in (1, 2, 4, 8):
        shards = shard_tracks(off, 24, 12, world)
        assert shards[0][0] == 0 and shards[-1][1] == 500
        assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
        per = [int(counts[lo:hi].sum()) for lo, hi, _ in shards]
        assert sum(per) == counts.sum()
        assert [s[2] for s in shards] == [int(counts[:lo].sum()) for lo, _, _ in shards]      # global index of the first window
        assert max(per) - min(per) <= counts.max() * 2                                         # cut at track boundaries only
    # degenerate: fewer tracks than ranks, and no windows at all
    assert len(shard_tracks(np.array([0, 100]), 24, 12, 4)) == 4
    assert shard_tracks(np.array([0, 5, 9]), 24, 12, 2) == [(0, 0, 0), (0, 2, 0)] or shard_tracks(np.array([0, 5, 9]), 24, 12, 2)[-1][1] == 2


def _ragged_worker(rank: int, world: int, port: int):
    from shopformer_b200.sharding import gather_ragged
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = torch.arange(5 + 3 * rank, dtype=torch.float32) + 100 * rank       # rank 0: 5 scores, rank 1: 8 scores
        allv, counts = gather_ragged(local)
        assert counts == [5, 8]
        assert torch.equal(allv, torch.cat([torch.arange(5, dtype=torch.float32), torch.arange(8, dtype=torch.float32) + 100]))
    finally:
        dist.destroy_process_group()


def test_ragged_score_allgather_world2():
    mp.spawn(_ragged_worker, args=(2, _free_port()), nprocs=2, join=True)
