"""(needs the stamps compiled in: `make -C computer-vision-shoplifting-detection_b200/csrc EXTRA=-DSF_STAMPS` after touching the kernel sources;
the default build leaves them out because they cost the kernels 2-5 %)
Per-op timeline of CTA 0 of the bf16 transformer (debug aid): python profiles/xf_timing.py"""
import ctypes as C, sys, os
FULL = os.environ.get("XF_TIMING_FULL") is not None
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import numpy as np, torch
import bench
from shopformer_b200 import native as N
lib = N.load()
model = bench.build_model("A").cuda()
eng = model._sf_engine()
tok = torch.randn(65536, 3, 136, device="cuda")
eng.reconstruct_tokens(tok, precision="bf16"); torch.cuda.synchronize()
lib.sfdbg_transformer_timing(1, None, 0)
eng.reconstruct_tokens(tok, precision="bf16"); torch.cuda.synchronize()
buf = (C.c_longlong * 1024)()
lib.sfdbg_transformer_timing(0, buf, 1024)
a = np.array(buf[:]).reshape(-1, 2)
end = np.where(a[:, 0] == 9999)[0]
lo, hi = end[0] + 1, end[1]           # second tile of the CTA
names = {0: "op start", 1: "w ready+sync", 2: "issued", 3: "mma done", 4: "epi done", 5: "attn QK done", 6: "ln rstd", 7: "attn softmax done"}
t0 = a[lo, 1]; prev = t0
agg = {}
for i in range(lo, hi + 1):
    op, ph = divmod(int(a[i, 0]), 8) if a[i, 0] != 9999 else (99, 5)
    dt = a[i, 1] - prev; prev = a[i, 1]
    agg[ph] = agg.get(ph, 0) + dt
    if FULL or 3 <= op < 8 or op == 99: print(f"op {op:2d} {names.get(ph, 'tile end'):14s} +{dt:6d}  @{a[i,1]-t0:7d}")
print({names.get(k, 'post/tile end'): int(v) for k, v in agg.items()}, "total", int(a[hi, 1] - t0))
