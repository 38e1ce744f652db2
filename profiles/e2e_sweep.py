"""End-to-end (host buffers) throughput of sf_runner_score / sf_runner_score_tracks against the chunk size.
    python profiles/e2e_sweep.py"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import numpy as np, torch
import bench
from shopformer_b200.engine import PackedTracks
from shopformer_b200.synthetic import synth_tracks, synth_windows
model = bench.build_model("A").cuda()
eng = model._sf_engine()
n = 65536
xs = torch.from_numpy(synth_windows(n, 24, 17, seed=1)[0]).pin_memory().numpy()
tr = synth_tracks(int(n / ((1515 - 24) / 12)), seed=4321)
kp2 = torch.from_numpy(np.ascontiguousarray(tr["kp"][:, :, :2])).pin_memory().numpy()
pin = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).pin_memory().numpy()
host = PackedTracks(kp=kp2, frame_no=pin(tr["frame_no"]), track_offsets=tr["track_offsets"], track_video=tr["track_video"], gt=pin(tr["gt"]), gt_offsets=tr["gt_offsets"])
for chunk in (4096, 8192, 16384, 32768, 65536):
    for _ in range(2):
        eng.score_host(xs, precision="tc", chunk=chunk)
    t0 = time.perf_counter()
    for _ in range(10):
        eng.score_host(xs, precision="tc", chunk=chunk)
    a = n * 10 / (time.perf_counter() - t0)
    for _ in range(2):
        r = eng.score_tracks_host(host, 24, 12, add_neck=False, precision="tc", chunk=chunk)
    t0 = time.perf_counter()
    for _ in range(10):
        r = eng.score_tracks_host(host, 24, 12, add_neck=False, precision="tc", chunk=chunk)
    b = r["n_windows"] * 10 / (time.perf_counter() - t0)
    print(f"chunk {chunk:6d}: pre-cut windows {a / 1e6:6.2f} M/s   from tracks {b / 1e6:6.2f} M/s", flush=True)
