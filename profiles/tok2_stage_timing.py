"""(needs the stamps compiled in: `make -C computer-vision-shoplifting-detection_b200/csrc EXTRA=-DSF_STAMPS` after touching the kernel sources;
the default build leaves them out because they cost the kernels 2-5 %)
Per-stage phases of tokenizer v2's epilogue teams (CTA 0, steady-state tile): python profiles/tok2_stage_timing.py [config]
For every stage of warp 4 (team 0) and warp 12 (team 1): table entry loaded -> waits done (start) -> body done -> arrived.
Goes with profiles/tok2_describe.py (what each stage / group is).  Needs the fine stamps compiled in:
`make -C computer-vision-shoplifting-detection_b200/csrc EXTRA=-DSF_TOK2_FINE_STAMPS` after touching tokenizer2_bf16.cu (they cost registers, so the
default build leaves them out)."""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import numpy as np, torch
import bench
from shopformer_b200 import native as N
from shopformer_b200 import configs as CFG
from shopformer_b200.synthetic import synth_windows
name = sys.argv[1] if len(sys.argv) > 1 else "A"
lib = N.load()
model = bench.build_model(name).cuda()
eng = model._sf_engine()
_, T, V = CFG.input_shape(name)
x = torch.from_numpy(synth_windows(65536, T, V, seed=1)[0]).cuda()
eng.tokenize(x, precision="bf16"); torch.cuda.synchronize()
lib.sfdbg_tokenizer2_timing(1, None, 0)
eng.tokenize(x, precision="bf16"); torch.cuda.synchronize()
buf = (C.c_longlong * 4096)()
lib.sfdbg_tokenizer2_timing(0, buf, 4096)
a = np.array(buf[:]).reshape(-1, 2)
g = a[:512]
t0 = int(g[(g[:, 0] >= 1000) & (g[:, 0] < 2000)][0, 1])
gi = {int(i) - 1000: int(t) - t0 for i, t in g if 1000 <= i < 2000}
gd = {int(i) - 3000: int(t) - t0 for i, t in g if 3000 <= i < 4000}
print("G groups: issue start / issue end")
for k in sorted(gi):
    print(f"  G{k:<3d} {gi[k]:7d} {gd.get(k, 0):7d}")
for k, reg in enumerate([a[512:1280], a[1280:]]):
    ph = {}
    for i, t in reg:
        i = int(i)
        if 2000 <= i < 10000:
            ph.setdefault(i % 1000, {})[i // 1000] = int(t) - t0
    print(f"team {k}: stage  loaded   start(waits done)  body done  arrived |  wait  body  fence+arrive")
    for e in sorted(ph):
        p = ph[e]
        l, s, b, r = p.get(5), p.get(2), p.get(6), p.get(7)
        if None in (l, s, b, r):
            continue
        cv = f"   CVT: first TMEM wait +{p[8]-s}, first pair stored +{p[9]-s}" if 8 in p and 9 in p else ""
        print(f"  E{k}.{e:<3d} {l:7d} {s:7d} {b:7d} {r:7d} | {s-l:5d} {b-s:5d} {r-b:5d}{cv}")
