"""Group / stage start stamps of CTA 0's second tile of tokenizer v2 (debug aid): python profiles/tok2_timing.py
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
raw = [a[512:1280], a[1280:]]
teams = [t[(t[:, 0] >= 2000) & (t[:, 0] < 3000)] for t in raw]
sub = [{k: {int(i) - k: int(t) for i, t in r[(r[:, 0] >= k) & (r[:, 0] < k + 500)]} for k in (1500, 3000, 5000, 6000)} for r in raw]
if len(g) == 0:
    raise SystemExit("no stamps (tokenizer v2 not used for this shape?)")
t0 = min([g[0, 1]] + [t[0, 1] for t in teams if len(t)])
print("MMA groups (id, start, delta to previous):")
prev = g[0, 1]
ga = a[:512]
gsub = {k: {int(i) - k: int(t) for i, t in ga[(ga[:, 0] >= k) & (ga[:, 0] < k + 500)]} for k in (7000, 8000, 9000)}
for i, t in g:
    j = int(i) - 1000
    det = ""
    if j in gsub[7000]: det += f"  waits+fence {t - gsub[7000][j]:5d}"
    if j in gsub[8000]: det += f"  issue {gsub[8000][j] - t:5d}"
    if j in gsub[9000] and j in gsub[8000]: det += f"  commit {gsub[9000][j] - gsub[8000][j]:5d}"
    print(f"  G{j:3d}  @{t-t0:7d}  +{t-prev:6d}{det}")
    prev = t
for k, e in enumerate(teams):
    print(f"epilogue team {k} stages (first warp of the team):")
    prev = e[0, 1]
    for i, t in e:
        j = int(i) - 2000
        d = sub[k]
        det = ""
        if j in d[1500]: det += f"  waited {t - d[1500][j]:5d}"
        if j in d[3000]: det += f"  g0-body {d[3000][j] - t:5d}"
        if j in d[5000]: det += f"  to-fence-done {d[5000][j] - t:5d}"
        if j in d[6000]: det += f"  to-arrived {d[6000][j] - t:5d}"
        print(f"  E{k}.{j:<3d}  @{t-t0:7d}  +{t-prev:6d}{det}")
        prev = t
print("tile span (G first -> last stamp):", g[-1, 1] - g[0, 1])
