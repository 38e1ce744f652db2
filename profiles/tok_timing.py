"""(needs the stamps compiled in: `make -C computer-vision-shoplifting-detection_b200/csrc EXTRA=-DSF_STAMPS` after touching the kernel sources;
the default build leaves them out because they cost the kernels 2-5 %)
Phase timeline of CTA 0 of the bf16 tokenizer (debug aid): python profiles/tok_timing.py"""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import numpy as np, torch
import bench
from shopformer_b200 import native as N
from shopformer_b200.synthetic import synth_windows
lib = N.load()
CFGN = sys.argv[1] if len(sys.argv) > 1 else "A"
from shopformer_b200 import configs as CFG
_, T_, V_ = CFG.input_shape(CFGN)
model = bench.build_model(CFGN).cuda()
eng = model._sf_engine()
x = torch.from_numpy(synth_windows(65536, T_, V_, seed=1)[0]).cuda()
eng.tokenize(x, precision="bf16"); torch.cuda.synchronize()
lib.sfdbg_tokenizer_timing(1, None, 0)
eng.tokenize(x, precision="bf16"); torch.cuda.synchronize()
buf = (C.c_longlong * 512)()
lib.sfdbg_tokenizer_timing(0, buf, 512)
a = np.array(buf[:]).reshape(-1, 2)
a = a[a[:, 0] >= 100]
names = {100: "group start", 101: "x0 affine done", 102: "m0 mix done"}
for b in range(4):
    names.update({110 + 10 * b: f"b{b} start", 111 + 10 * b: f"b{b} Q-epi done", 112 + 10 * b: f"b{b} zero A_big + sync", 200 + 10 * b: f"b{b} GEMM2 done",
                  113 + 10 * b: f"b{b} P mma done", 114 + 10 * b: f"b{b} g-epi done", 115 + 10 * b: f"b{b} sync", 116 + 10 * b: f"b{b} conv mma done",
                  117 + 10 * b: f"b{b} x-epi done", 118 + 10 * b: f"b{b} conv issued", 119 + 10 * b: f"b{b} conv mbar"})
# print the 3rd window of the CTA (steady state)
starts = np.where(a[:, 0] == 100)[0]
lo, hi = starts[2], starts[3]
t0 = a[lo, 1]
prev = t0
for i in range(lo, hi + 1):
    print(f"{names.get(int(a[i,0]), a[i,0]):24s} +{a[i,1]-prev:7d}  @{a[i,1]-t0:7d}")
    prev = a[i, 1]
