"""Probe: tokenizer TMA reading poses straight from pinned host memory (no H2D copy): python profiles/zero_copy_probe.py"""
import ctypes as C, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import numpy as np, torch
import bench
from shopformer_b200 import native as N
from shopformer_b200.synthetic import synth_windows
lib = N.load()
model = bench.build_model("A").cuda()
eng = model._sf_engine()
n = 65536
xs = synth_windows(n, 24, 17, seed=1)[0]
xp = torch.from_numpy(xs).pin_memory()
xd = xp.cuda()
ref = eng.score_windows(xd, precision="bf16")
ws_bytes = lib.sf_workspace_bytes(eng._h, n, 24)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
out = torch.empty(n, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run(ptr):
    rc = lib.sf_score_windows(eng._h, C.c_void_p(ptr), n, 24, N.SF_REDUCE_MEAN, N.SF_PREC_BF16, C.c_void_p(out.data_ptr()), None, None,
                              C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(st))
    assert rc == 0, lib.sf_last_error()
for name, ptr in (("device", xd.data_ptr()), ("pinned host (zero copy)", xp.data_ptr())):
    for _ in range(3): run(ptr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run(ptr)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:26s} {e0.elapsed_time(e1)/10:.3f} ms/step   equal to device-resident result: {bool(torch.equal(out, ref))}")
import time
xn = xp.numpy()
for name, arr in (("score_host pinned", xn), ("score_host pageable", xs)):
    eng.score_host(arr, precision="bf16", chunk=16384)
    t0 = time.perf_counter()
    for _ in range(5): h = eng.score_host(arr, precision="bf16", chunk=16384)
    dt = (time.perf_counter() - t0) / 5
    print(f"{name:26s} {dt*1e3:.3f} ms/step  equal: {bool(np.array_equal(h, ref.cpu().numpy()))}")
