"""(needs the stamps compiled in: `make -C computer-vision-shoplifting-detection_b200/csrc EXTRA=-DSF_STAMPS` after touching the kernel sources;
the default build leaves them out because they cost the kernels 2-5 %)
Group / stage start stamps of CTA 0's second tile of tokenizer v2 (debug aid): python profiles/tok2_timing.py
MMA warp: one stamp per G group just before its MMAs are issued (id 1000 + g); epilogue warp 4: one stamp per E stage
after its waits (id 2000 + e).  Prints both timelines relative to the first stamp, cycles @ SM clock."""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import numpy as np, torch
import bench
from shopformer_b200 import native as N
from shopformer_b200.synthetic import synth_windows
lib = N.load()
model = bench.build_model("A").cuda()
eng = model._sf_engine()
x = torch.from_numpy(synth_windows(65536, 24, 17, seed=1)[0]).cuda()
eng.tokenize(x, precision="bf16"); torch.cuda.synchronize()
lib.sfdbg_tokenizer2_timing(1, None, 0)
eng.tokenize(x, precision="bf16"); torch.cuda.synchronize()
buf = (C.c_longlong * 4096)()
lib.sfdbg_tokenizer2_timing(0, buf, 4096)
a = np.array(buf[:]).reshape(-1, 2)
g = a[:512][(a[:512, 0] >= 1000) & (a[:512, 0] < 2000)]
teams = [a[512:1280], a[1280:]]
teams = [t[(t[:, 0] >= 2000) & (t[:, 0] < 3000)] for t in teams]
if len(g) == 0:
    raise SystemExit("no stamps (tokenizer v2 not used for this shape?)")
t0 = min([g[0, 1]] + [t[0, 1] for t in teams if len(t)])
done = {int(i) - 3000: int(t) for i, t in a[:512] if 3000 <= i < 4000}
print("MMA groups (id, issue start, delta to previous start, issue duration = start -> after commit, then wait for the next group's dependencies):")
prev = g[0, 1]
for k, (i, t) in enumerate(g):
    gi = int(i) - 1000
    dur = done.get(gi, t) - t
    nxt = g[k + 1, 1] - done.get(gi, t) if k + 1 < len(g) and gi in done else 0
    print(f"  G{gi:3d}  @{t-t0:7d}  +{t-prev:6d}   issue {dur:5d}   wait {nxt:5d}")
    prev = t
for k, e in enumerate(teams):
    print(f"epilogue team {k} stages (first warp of the team):")
    prev = e[0, 1]
    for i, t in e:
        print(f"  E{k}.{i-2000:<3d}  @{t-t0:7d}  +{t-prev:6d}")
        prev = t
print("tile span (G first -> last stamp):", g[-1, 1] - g[0, 1])
