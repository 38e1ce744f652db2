"""ncu target: a few launches of the scoring path of one config (default B) on 65,536 windows.
    python profiles/r2_once_cfg.py [B|C|A1]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import torch
import bench
from shopformer_b200 import configs as CFG
from shopformer_b200.synthetic import synth_windows
cfg = sys.argv[1] if len(sys.argv) > 1 else "B"
_, T, V = CFG.input_shape(cfg)
model = bench.build_model(cfg).cuda()
eng = model._sf_engine()
x = torch.from_numpy(synth_windows(65536, T, V, seed=1)[0]).cuda()
for _ in range(3):
    s = eng.score_windows(x, precision="tc")
torch.cuda.synchronize()
print(cfg, "ok", float(s.mean()), eng.tc_formats(T))
