"""ncu target of round 2: a few launches of the scoring path (tokenizer v2 + tensor-core transformer) on 65,536 config-A
windows, then of the HBM-bound kernels (windowing from packed tracks, stand-alone score) on ~1 M windows.
    python profiles/r2_once.py [path|windowing|both]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import torch
import bench
from shopformer_b200.engine import DeviceTracks, window_normalize
from shopformer_b200.synthetic import synth_tracks_device, synth_windows
what = sys.argv[1] if len(sys.argv) > 1 else "both"
dev = torch.device("cuda", 0)
model = bench.build_model("A").cuda()
eng = model._sf_engine()
if what in ("path", "both"):
    x = torch.from_numpy(synth_windows(65536, 24, 17, seed=1)[0]).cuda()
    for _ in range(3):
        s = eng.score_windows(x, precision="tc")
    torch.cuda.synchronize()
    print("path ok", float(s.mean()), eng.tc_formats(24))
if what in ("windowing", "both"):
    n_tracks = 1_048_576 // ((1512 - 24) // 12 + 1)
    kp, fno, off, vid, gt, gto = synth_tracks_device(n_tracks, 1512, dev, seed=1234, channels=3)
    dt = DeviceTracks.__new__(DeviceTracks)
    dt.host, dt.device, dt.kp, dt.frame_no, dt.gt = None, dev, kp, fno, gt
    dt._off, dt._vid, dt._gto = off, vid, gto
    for _ in range(2):
        out = window_normalize(dt, 24, 12, num_keypoints=17, add_neck=False)
    n = out["n_windows"]
    tok = torch.randn(n, 3, 136, device=dev)
    rec = torch.randn(n, 3, 136, device=dev)
    for _ in range(2):
        sc = eng.normality_score(tok, rec)
    torch.cuda.synchronize()
    print("windowing ok", n, float(sc.mean()))
