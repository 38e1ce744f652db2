"""GPU: the rows either side of the scoring path -- packed tracks -> scores in one call (device and host tracks), and the
on-device post-processing (per-video aggregation, AUC-ROC / AUC-PR / thresholding) against numpy / scikit-learn, which
is what the reference calls (shopformer_2/evaluate.py:65-118, shopformer_2/utils/metrics.py:21-188)."""
import numpy as np
import pytest
import torch

from helpers import build_model
from shopformer_b200.engine import DeviceTracks, PackedTracks, ranking_metrics, video_aggregate, window_normalize
from shopformer_b200.synthetic import synth_tracks

pytestmark = pytest.mark.gpu


def packed(n_tracks, seed, **kw):
    tr = synth_tracks(n_tracks, seed=seed, **kw)
    return PackedTracks(kp=tr["kp"], frame_no=tr["frame_no"], track_offsets=tr["track_offsets"], track_video=tr["track_video"],
                        gt=tr["gt"], gt_offsets=tr["gt_offsets"])


@pytest.fixture(scope="module")
def eng_a(dropin1, dropin2):
    m = build_model(dropin1, dropin2, "A").cuda()
    eng = m._sf_engine()
    eng._keepalive = m
    return eng


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_score_from_tracks_equals_window_then_score(eng_a, precision):
    tracks = packed(60, 11, min_len=10, max_len=600, gap_every=120)
    dev = DeviceTracks(tracks, torch.device("cuda"))
    win = window_normalize(dev, 24, 12, num_keypoints=17, add_neck=False)
    want = eng_a.score_windows(win["poses"], precision=precision)
    got = eng_a.score_tracks(dev, 24, 12, add_neck=False, precision=precision)
    assert got["n_windows"] == win["n_windows"] > 100
    assert torch.equal(got["labels"], win["labels"])
    assert torch.equal(got["window_track"], win["window_track"]) and torch.equal(got["window_start"], win["window_start"])
    assert torch.equal(got["scores"], want)                      # same windows through the same kernels: same bits


def test_score_from_tracks_several_passes(eng_a):
    """More windows than one internal pass (131,072): the pass boundary must not show."""
    tracks = packed(1200, 3, min_len=1000, max_len=2200, gap_every=700)
    dev = DeviceTracks(tracks, torch.device("cuda"))
    got = eng_a.score_tracks(dev, 24, 12, add_neck=False, precision="tc")
    n = got["n_windows"]
    assert n > 131072 + 1000
    win = window_normalize(dev, 24, 12, num_keypoints=17, add_neck=False)
    assert win["n_windows"] == n
    idx = torch.cat([torch.arange(0, 4096), torch.arange(131072 - 2048, 131072 + 2048), torch.arange(n - 4096, n)]).cuda()
    want = eng_a.score_windows(win["poses"][idx], precision="tc")
    assert torch.allclose(got["scores"][idx], want, rtol=1e-6, atol=0)
    assert torch.isfinite(got["scores"]).all()


@pytest.mark.parametrize("channels", [3, 2])
def test_host_tracks_runner_matches_device_tracks(eng_a, channels):
    tracks = packed(300, 21, min_len=5, max_len=900, gap_every=200)
    dev = DeviceTracks(tracks, torch.device("cuda"))
    want = eng_a.score_tracks(dev, 24, 12, add_neck=False, precision="tc")
    host = tracks
    if channels == 2:                                            # ingest dropped the confidence: a third fewer bytes to upload
        host = PackedTracks(kp=np.ascontiguousarray(tracks.kp[:, :, :2]), frame_no=tracks.frame_no, track_offsets=tracks.track_offsets,
                            track_video=tracks.track_video, gt=tracks.gt, gt_offsets=tracks.gt_offsets)
    for chunk in (4096, 1000):                                   # several groups of whole tracks, ragged last group
        got = eng_a.score_tracks_host(host, 24, 12, add_neck=False, precision="tc", chunk=chunk)
        assert got["n_windows"] == want["n_windows"]
        assert np.array_equal(got["labels"], want["labels"].cpu().numpy())
        assert np.array_equal(got["window_track"], want["window_track"].cpu().numpy())
        assert np.array_equal(got["window_start"], want["window_start"].cpu().numpy())
        assert np.allclose(got["scores"], want["scores"].cpu().numpy(), rtol=1e-6, atol=0)


def test_host_tracks_runner_edge_cases(eng_a):
    # no track long enough for a window
    tr = packed(5, 1, min_len=3, max_len=20, gap_every=0)
    got = eng_a.score_tracks_host(tr, 24, 12, add_neck=False, precision="tc")
    assert got["n_windows"] == 0 and got["scores"].shape == (0,)
    # a single long track, no ground truth
    tr = packed(1, 2, min_len=800, max_len=800, gap_every=0)
    tr.gt = None
    tr.gt_offsets = None
    got = eng_a.score_tracks_host(tr, 24, 12, add_neck=False, precision="tc", chunk=16)
    assert got["n_windows"] > 40 and (got["labels"] == 0).all() and np.isfinite(got["scores"]).all()


def test_video_aggregate_matches_numpy():
    rs = np.random.RandomState(0)
    n, nv = 50_000, 37
    scores = rs.gamma(2.0, 1.0, n).astype(np.float32)
    scores[rs.randint(0, n, 500)] = scores[0]                    # ties
    vid = rs.randint(0, nv - 2, n).astype(np.int32)              # the last two videos have no windows
    vid[:5] = [nv + 3, -1, nv, 0, 0]                              # out-of-range ids are ignored
    labels = rs.randint(0, 2, n).astype(np.int32)
    out = video_aggregate(torch.from_numpy(scores).cuda(), torch.from_numpy(vid).cuda(), torch.from_numpy(labels).cuda(), nv)
    for v in range(nv):
        sel = np.flatnonzero(vid == v)
        assert int(out["count"][v]) == len(sel)
        if len(sel) == 0:
            assert np.isnan(float(out["max"][v])) and int(out["label"][v]) == 0
            continue
        s = np.array([float(x) for x in scores[sel]])            # the reference appends python floats
        assert float(out["max"][v]) == np.max(s)
        assert abs(float(out["mean"][v]) - np.mean(s)) <= 1e-12 * abs(np.mean(s))
        assert abs(float(out["percentile_95"][v]) - np.percentile(s, 95)) <= 1e-12 * abs(np.percentile(s, 95))
        assert int(out["label"][v]) == labels[sel[-1]]            # `video_labels[video_id] = info['label']`: the last one wins


@pytest.mark.parametrize("n,ties", [(257, False), (10_000, True), (300_000, True)])
def test_ranking_metrics_match_sklearn(n, ties, dropin2):
    from sklearn.metrics import average_precision_score, roc_auc_score
    rs = np.random.RandomState(n)
    labels = (rs.random_sample(n) < 0.3).astype(np.int64)
    scores = (rs.randn(n) + 0.8 * labels).astype(np.float32)
    if ties:
        scores = np.round(scores * 50) / 50                      # heavy ties: every tie group mixes both classes
    ref = dropin2["metrics"].compute_metrics(labels, scores)     # the reference's entry point (sklearn on the host)
    got = ranking_metrics(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda())
    assert abs(got["auc_roc"] - roc_auc_score(labels, scores)) < 1e-12
    assert abs(got["auc_pr"] - average_precision_score(labels, scores)) < 1e-12
    assert abs(got["auc_roc"] - ref["auc_roc"]) < 1e-12 and abs(got["auc_pr"] - ref["auc_pr"]) < 1e-12
    assert got["threshold"] == pytest.approx(float(ref["threshold"]), abs=0)
    for k in ("accuracy", "precision", "recall", "f1"):
        assert abs(got[k] - ref[k]) < 1e-12, k
    # explicit threshold
    got = ranking_metrics(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(), threshold=0.5)
    pred = scores >= np.float32(0.5)
    assert got["tp"] == int((pred & (labels == 1)).sum()) and got["fp"] == int((pred & (labels == 0)).sum())
    assert got["tn"] == int((~pred & (labels == 0)).sum()) and got["fn"] == int((~pred & (labels == 1)).sum())


def test_ranking_metrics_degenerate_labels():
    s = torch.rand(1000, device="cuda")
    one = ranking_metrics(s, torch.zeros(1000, dtype=torch.int32, device="cuda"))
    assert one["auc_roc"] == 0.5 and one["auc_pr"] == 0.0          # the reference's except-ValueError conventions
    allpos = ranking_metrics(s, torch.ones(1000, dtype=torch.int32, device="cuda"))
    assert allpos["auc_roc"] == 0.5 and allpos["auc_pr"] == pytest.approx(1.0)


def test_sweep_end_to_end_on_device(eng_a):
    """tracks -> scores -> per-video max -> video-level AUC without the scores leaving the device until the metrics."""
    from sklearn.metrics import roc_auc_score
    tracks = packed(200, 8, min_len=100, max_len=700, gap_every=250)
    dev = DeviceTracks(tracks, torch.device("cuda"))
    out = eng_a.score_tracks(dev, 24, 12, add_neck=False, precision="tc")
    vid_of_track = torch.from_numpy(np.asarray(tracks.track_video, dtype=np.int32)).cuda()
    vid = vid_of_track[out["window_track"].long()]
    nv = int(tracks.track_video.max()) + 1
    agg = video_aggregate(out["scores"], vid, out["labels"], nv)
    has = agg["count"] > 0
    got = ranking_metrics(agg["max"][has].float(), agg["label"][has])
    s = out["scores"].cpu().numpy()
    v = vid.cpu().numpy()
    lab = out["labels"].cpu().numpy()
    vmax = np.array([s[v == k].max() for k in range(nv) if (v == k).any()], dtype=np.float32)
    vlab = np.array([lab[np.flatnonzero(v == k)[-1]] for k in range(nv) if (v == k).any()])
    if 0 < vlab.sum() < len(vlab):
        assert abs(got["auc_roc"] - roc_auc_score(vlab, vmax)) < 1e-12
