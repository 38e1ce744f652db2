"""GPU: streaming mode (BASELINE config #5, scaled down) -- rings + pre-cut normalisation + scoring per tick,
eager and CUDA-graph replay, against oracle windows cut from the same detections."""
import numpy as np
import pytest
import torch

import oracle.scoring_oracle as O
import oracle.windowing_oracle as W
from helpers import build_model, oracle_kwargs, rel_err
from shopformer_b200.streaming import StreamScorer

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_graph,precision", [(False, "fp32"), (True, "fp32"), (True, "bf16")])
def test_streaming_ticks_match_oracle(use_graph, precision, dropin1, dropin2):
    n_streams, T, stride, n_ticks = 37, 24, 12, 5
    rs = np.random.RandomState(3)
    base = rs.uniform(100, 900, (n_streams, 1, 17, 2))
    frames = base + np.cumsum(rs.randn(n_streams, T + stride * n_ticks, 17, 2) * 2.0, axis=1)
    frames = np.concatenate([frames, rs.uniform(0.1, 1, frames.shape[:3] + (1,))], axis=-1).astype(np.float32)
    frames[3, 30:33, 5] = 0.0                                  # a few missing keypoints
    model = build_model(dropin1, dropin2, "A")
    cpu_sd = {k: v.clone() for k, v in model.state_dict().items()}
    kw = oracle_kwargs(model, "A")
    model = model.cuda()
    sc = StreamScorer(model._sf_engine(), n_streams, T, stride, precision=precision, use_graph=use_graph)
    pos = 0
    sc.tick(frames[:, pos:pos + stride]); pos += stride      # fill the rings (first window incomplete)
    for _ in range(n_ticks):
        got = sc.tick(frames[:, pos:pos + stride]); pos += stride
        wins = np.stack([W.normalize_window(frames[i, pos - T:pos, :, :2].astype(np.float64)) for i in range(n_streams)])
        x = torch.from_numpy(np.transpose(wins, (0, 3, 1, 2)).astype(np.float32))
        ref = O.score_windows(cpu_sd, x, dtype=torch.float64, **kw)["score"].numpy()
        assert rel_err(got, ref) < (5e-5 if precision == "fp32" else 1e-2)
    assert sc.ticks == n_ticks + 1
