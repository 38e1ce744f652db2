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
