"""CPU: the oracle restatement against the golden vectors produced by the reference itself."""
import json

import numpy as np
import pytest
import torch

import oracle.scoring_oracle as O
import oracle.windowing_oracle as W
from helpers import build_model, max_abs_rel, oracle_kwargs, rel_err
from shopformer_b200 import configs as CFG
from shopformer_b200.synthetic import synth_poselift_video


@pytest.mark.parametrize("name", CFG.ALL_CONFIGS)
def test_scoring_oracle_matches_reference(name, golden_dir, dropin1, dropin2):
    g = np.load(golden_dir / f"score_{name}.npz")
    model = build_model(dropin1, dropin2, name)
    kw = oracle_kwargs(model, name)
    assert list(g["strides"]) == kw["strides"]
    assert int(g["nhead"]) == kw["nhead"]
    assert int(g["pool_tokens"]) == (kw["pool_tokens"] or 0)
    x = torch.from_numpy(g["poses"])
    o64 = O.score_windows(model.state_dict(), x, dtype=torch.float64, **kw)
    assert rel_err(o64["score"].numpy(), g["score64"]) < 1e-11
    n = g["tokens64"].shape[0]
    assert max_abs_rel(o64["tokens"][:n].numpy(), g["tokens64"]) < 1e-11
    assert max_abs_rel(o64["recon"][:n].numpy(), g["recon64"]) < 1e-11
    o32 = O.score_windows(model.state_dict(), x, dtype=torch.float32, **kw)
    assert rel_err(o32["score"].numpy(), g["score64"]) < 2e-5      # fp32 arithmetic noise
    assert rel_err(g["score32"], g["score64"]) < 2e-5               # the reference's own fp32 noise
    if kw["variant"] == 2:
        per_tok = O.score_v2(o64["tokens"], o64["recon"], "none").numpy()
        assert rel_err(per_tok, g["per_token64"]) < 1e-10
        with pytest.raises(ValueError):
            O.score_v2(o64["tokens"], o64["recon"], "sum")


def test_oracle_accepts_btvc_layout(golden_dir, dropin1, dropin2):
    g = np.load(golden_dir / "score_A.npz")
    model = build_model(dropin1, dropin2, "A")
    kw = oracle_kwargs(model, "A")
    x = torch.from_numpy(g["poses"][:4])
    a = O.tokenize(model.state_dict(), x.double(), kw["strides"])
    b = O.tokenize(model.state_dict(), x.double().permute(0, 2, 3, 1).contiguous(), kw["strides"])
    assert torch.equal(a, b)


def test_stride_rules():
    assert O.strides_v1(24, 2) == [2, 2, 2, 1]
    assert O.strides_v1(12, 2) == [2, 2, 1, 1]
    assert O.strides_v2(12, 2) == [3, 2, 1, 1]
    assert O.strides_v2(24, 2) == [3, 2, 2, 1]
    assert O.v2_needs_pool(24, 5, O.strides_v2(24, 5)) and not O.v2_needs_pool(24, 2, O.strides_v2(24, 2))
    assert [O.conv_len(24, s) for s in (1, 2, 3)] == [24, 12, 8]


WIN_CASES = [("v1_T12", 1, dict(seq_len=12, stride=6, num_keypoints=17)),
             ("v1_T24", 1, dict(seq_len=24, stride=12, num_keypoints=17)),
             ("v1_T24_nonorm", 1, dict(seq_len=24, stride=12, num_keypoints=17, normalize=False)),
             ("v1_T12_V18", 1, dict(seq_len=12, stride=6, num_keypoints=18)),          # variant 1: zero 18th keypoint, no neck
             ("v1_T12_conf", 1, dict(seq_len=12, stride=6, num_keypoints=17, include_confidence=True)),
             ("v2_T24", 2, dict(seq_len=24, stride=12, num_keypoints=17)),
             ("v2_T12_neck", 2, dict(seq_len=12, stride=6, num_keypoints=18)),
             ("v2_T12_gap9", 2, dict(seq_len=12, stride=5, num_keypoints=18, max_gap=9))]


def fixture_videos():
    return {f"vid{n:02d}": synth_poselift_video(seed=100 + n, n_frames=300 + 40 * n) for n in range(2)}


@pytest.mark.parametrize("tag,variant,kw", WIN_CASES)
def test_windowing_oracle_bit_exact(tag, variant, kw, golden_dir):
    g = np.load(golden_dir / "windowing.npz")
    wins, labels, fidx = [], [], []
    for name, (frames, gt) in sorted(fixture_videos().items()):
        okw = {k: v for k, v in kw.items() if k != "include_confidence"}
        w, l, f = W.extract_windows(frames, gt, variant=variant, channels=3 if kw.get("include_confidence") else 2, **okw)
        wins += w; labels += l; fidx += f
    assert np.array_equal(np.asarray(labels), g[f"{tag}_labels"])
    assert np.array_equal(np.asarray(fidx), g[f"{tag}_frame_indices"])
    assert np.array_equal(np.stack(wins), g[f"{tag}_windows"])            # bit-exact, floats included
    assert W.windows_as_model_input(wins).shape == (len(wins), 3 if kw.get("include_confidence") else 2, kw["seq_len"], kw["num_keypoints"])


def test_windowing_oracle_edge_cases():
    # empty input, a track shorter than T, an all-zero window, all-invalid normalisation
    assert W.extract_windows({}, None, seq_len=12, stride=6) == ([], [], [])
    short = {f: {1: [None, np.ones((17, 3))]} for f in range(5)}
    assert W.extract_windows(short, None, seq_len=12, stride=6)[0] == []
    zeros = {f: {1: [None, np.zeros((17, 3))]} for f in range(12)}
    w, l, _ = W.extract_windows(zeros, None, seq_len=12, stride=6)
    assert len(w) == 1 and not w[0].any() and l == [0]
    one = np.zeros((12, 17, 2)); one[3, 4] = (5.0, 7.0)
    n = W.normalize_window(one)
    assert np.allclose(n[3, 4], 0.0) and np.allclose(n[0, 0], (-5e6, -7e6))  # scale = 0 + 1e-6, invalid rows still move


def test_manifest_matches_dropin(golden_dir, dropin1, dropin2):
    man = json.load(open(golden_dir / "manifest.json"))
    for name in CFG.ALL_CONFIGS:
        model = build_model(dropin1, dropin2, name)
        sd = model.state_dict()
        assert list(sd.keys()) == man["configs"][name]["keys"]
        assert [list(v.shape) for v in sd.values()] == man["configs"][name]["shapes"]
        assert sum(p.numel() for p in model.parameters()) == man["configs"][name]["params"]
