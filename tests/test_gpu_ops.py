"""GPU: the scoring path as ``torch.ops.shopformer_b200.*`` (BASELINE north star: "drop-in PyTorch custom ops"): the ops
the facades call reach the same native kernels as the engine, have shape functions (fake tensors), and are CUDA-only."""
import numpy as np
import pytest
import torch

from helpers import build_model
from shopformer_b200 import ops
from shopformer_b200.synthetic import synth_windows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model_a(dropin1, dropin2):
    return build_model(dropin1, dropin2, "A").cuda()


def test_ops_match_engine_calls(model_a):
    eng = model_a._sf_engine()
    h = ops.model_handle(eng)
    x = torch.from_numpy(synth_windows(64, 24, 17, seed=2)[0]).cuda()
    ns = torch.ops.shopformer_b200
    for prec in ("fp32", "tc"):
        tok = ns.tokenize(x, h, prec)
        assert torch.equal(tok, eng.tokenize(x, precision=prec))
        rec = ns.reconstruct_tokens(tok, h, prec)
        assert torch.equal(rec, eng.reconstruct_tokens(tok, precision=prec))
        assert torch.equal(ns.normality_score(tok, rec, h, "mean"), eng.normality_score(tok, rec))
        s = ns.score_fused(x, h, "mean", prec)
        assert torch.equal(s, eng.score_windows(x, precision=prec))
        s3, t3, r3 = ns.score_fused_full(x, h, prec)
        assert torch.equal(s3, s) and torch.equal(t3, tok) and torch.equal(r3, rec)
    # the facade's forward goes through the ops: same numbers
    with torch.no_grad():
        out = model_a(x)
    assert torch.equal(out["normality_score"], ns.score_fused(x, h, "mean", "tc"))


def test_window_normalize_op_matches_reference_maths():
    rs = np.random.RandomState(0)
    raw = rs.uniform(100, 900, (5, 24, 17, 3)).astype(np.float32)
    raw[1, 3, 4, :2] = 0.0                                   # an undetected keypoint is excluded from the statistics
    got = torch.ops.shopformer_b200.window_normalize(torch.from_numpy(raw).cuda(), 17, True).cpu().numpy()
    xy = raw[..., :2].astype(np.float64)
    for b in range(5):
        valid = (xy[b, ..., 0] != 0) | (xy[b, ..., 1] != 0)
        c = xy[b][valid].mean(axis=0)
        sc = np.abs(xy[b][valid] - c).max() + 1e-6
        want = np.transpose((xy[b] - c) / sc, (2, 0, 1))
        assert np.max(np.abs(got[b] - want)) < 2e-6


def test_ops_have_shape_functions_and_are_cuda_only(model_a):
    from torch._subclasses.fake_tensor import FakeTensorMode
    eng = model_a._sf_engine()
    h = ops.model_handle(eng)
    with FakeTensorMode():
        x = torch.empty(7, 2, 24, 17, device="cuda")
        tok = torch.ops.shopformer_b200.tokenize(x, h, "tc")
        assert tuple(tok.shape) == (7, 3, 136)
        assert tuple(torch.ops.shopformer_b200.score_fused(x, h, "mean", "tc").shape) == (7,)
        assert tuple(torch.ops.shopformer_b200.window_normalize(torch.empty(4, 24, 17, 3, device="cuda"), 18, True).shape) == (4, 2, 24, 18)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.shopformer_b200.tokenize(torch.zeros(1, 2, 24, 17), h, "tc")        # no CPU kernel is registered
    with pytest.raises(RuntimeError):
        torch.ops.shopformer_b200.tokenize(torch.zeros(1, 2, 24, 17, device="cuda"), 12345, "tc")   # not a live handle
