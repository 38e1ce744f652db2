"""GPU parity of the windowing / normalisation kernels: integer outputs bit-exact against the
reference's golden fixture and the oracle, float outputs within fp32 rounding of the reference
(which normalises in float64 before casting)."""
import numpy as np
import pytest
import torch

import oracle.windowing_oracle as W
from shopformer_b200.engine import DeviceTracks, PackedTracks, window_normalize
from shopformer_b200.ingest import pack_videos
from shopformer_b200.synthetic import synth_poselift_video, synth_tracks
from test_oracle_golden import WIN_CASES, fixture_videos

pytestmark = pytest.mark.gpu
FLOAT_TOL = 2e-6     # absolute, on values normalised to [-1, 1]


def packed_fixture():
    vids = [(name, frames, gt) for name, (frames, gt) in sorted(fixture_videos().items())]
    return pack_videos(vids)


@pytest.mark.parametrize("tag,variant,kw", WIN_CASES)
def test_windowing_matches_reference_golden(tag, variant, kw, golden_dir):
    g = np.load(golden_dir / "windowing.npz")
    tracks = packed_fixture()
    dev = DeviceTracks(tracks, torch.device("cuda"))
    # variant 1 never synthesises a neck (an 18th keypoint is a zero row); variant 2 adds it iff V == 18
    out = window_normalize(dev, kw["seq_len"], kw["stride"], num_keypoints=kw["num_keypoints"], max_gap=kw.get("max_gap", 5),
                           normalize=kw.get("normalize", True), want_frame_indices=True,
                           add_neck=(variant == 2 and kw["num_keypoints"] == 18), include_confidence=kw.get("include_confidence", False))
    gold = g[f"{tag}_windows"]                                   # (N,T,V,C)
    assert out["n_windows"] == gold.shape[0]
    assert np.array_equal(out["labels"].cpu().numpy(), g[f"{tag}_labels"])              # bit-exact
    assert np.array_equal(out["frame_indices"].cpu().numpy(), g[f"{tag}_frame_indices"])  # bit-exact
    poses = out["poses"].cpu().numpy()                           # (N,2,T,V)
    want = np.transpose(gold, (0, 3, 1, 2))
    if kw.get("normalize", True):
        assert np.max(np.abs(poses - want)) < FLOAT_TOL
    else:
        assert np.array_equal(poses, want)                        # pure gather: bit-exact
    if variant == 2:
        vids = [tracks.video_names[tracks.track_video[t]] for t in out["window_track"].cpu().numpy()]
        assert vids == list(g[f"{tag}_video_ids"])


def test_windowing_matches_oracle_on_packed_tracks():
    """Bench-style synthetic tracks (dropped frames, >5-frame holes, zeroed keypoints)."""
    tr = synth_tracks(40, seed=5, min_len=20, max_len=400, gap_every=90)
    tracks = PackedTracks(kp=tr["kp"], frame_no=tr["frame_no"], track_offsets=tr["track_offsets"],
                          track_video=tr["track_video"], gt=tr["gt"], gt_offsets=tr["gt_offsets"])
    dev = DeviceTracks(tracks, torch.device("cuda"))
    for T, stride, V in ((24, 12, 17), (12, 6, 18), (24, 7, 17)):
        out = window_normalize(dev, T, stride, num_keypoints=V, want_frame_indices=True)
        wins, labels, fidx = [], [], []
        for i in range(tracks.n_tracks):
            a, b = tracks.track_offsets[i], tracks.track_offsets[i + 1]
            frames = {int(f): {0: [None, tracks.kp[j].astype(np.float64)]} for j, f in zip(range(a, b), tracks.frame_no[a:b])}
            g0, g1 = tracks.gt_offsets[i], tracks.gt_offsets[i + 1]
            w, l, f = W.extract_windows(frames, tracks.gt[g0:g1], seq_len=T, stride=stride, num_keypoints=V,
                                        variant=2 if V == 18 else 1)
            wins += w; labels += l; fidx += f
        assert out["n_windows"] == len(wins) > 0
        assert out["n_windows"] < out["capacity"]                 # some candidates were rejected for gaps
        assert np.array_equal(out["labels"].cpu().numpy(), np.asarray(labels))
        assert np.array_equal(out["frame_indices"].cpu().numpy(), np.asarray(fidx))
        want = np.transpose(np.stack(wins), (0, 3, 1, 2))
        assert np.max(np.abs(out["poses"].cpu().numpy() - want)) < FLOAT_TOL


def test_windowing_edge_cases():
    # no tracks at all / every track shorter than T -> zero windows, no kernel faults
    empty = PackedTracks(kp=np.zeros((0, 17, 3), np.float32), frame_no=np.zeros(0, np.int32),
                         track_offsets=np.zeros(1, np.int64), track_video=np.zeros(0, np.int32))
    out = window_normalize(DeviceTracks(empty, torch.device("cuda")), 24, 12)
    assert out["n_windows"] == 0 and out["poses"].shape == (0, 2, 24, 17)
    short = PackedTracks(kp=np.ones((10, 17, 3), np.float32), frame_no=np.arange(10, dtype=np.int32),
                         track_offsets=np.array([0, 10], np.int64), track_video=np.zeros(1, np.int32))
    assert window_normalize(DeviceTracks(short, torch.device("cuda")), 24, 12)["n_windows"] == 0
    # all-zero window: centre 0 / scale 1 branch -> zeros out, label 0 without GT
    z = PackedTracks(kp=np.zeros((24, 17, 3), np.float32), frame_no=np.arange(24, dtype=np.int32),
                     track_offsets=np.array([0, 24], np.int64), track_video=np.zeros(1, np.int32))
    out = window_normalize(DeviceTracks(z, torch.device("cuda")), 24, 12)
    assert out["n_windows"] == 1 and not out["poses"].any() and out["labels"].tolist() == [0]


def test_windows_feed_the_scorer(dropin1, dropin2, monkeypatch):
    """tracks -> windows -> scores entirely on the device equals oracle windows -> oracle scores."""
    import oracle.scoring_oracle as O
    from helpers import build_model, oracle_kwargs, rel_err
    monkeypatch.setenv("SHOPFORMER_B200_PRECISION", "fp32")          # checked at the fp32 kernels' 5e-5
    tracks = packed_fixture()
    out = window_normalize(DeviceTracks(tracks, torch.device("cuda")), 24, 12)
    model = build_model(dropin1, dropin2, "A")
    wins = []
    for name, (frames, gt) in sorted(fixture_videos().items()):
        wins += W.extract_windows(frames, gt, seq_len=24, stride=12)[0]
    ref = O.score_windows(model.state_dict(), torch.from_numpy(W.windows_as_model_input(wins)), dtype=torch.float64,
                          **oracle_kwargs(model, "A"))
    model = model.cuda()
    with torch.no_grad():
        s = model(out["poses"])["normality_score"]
    assert rel_err(s.cpu().numpy(), ref["score"].numpy()) < 5e-5
