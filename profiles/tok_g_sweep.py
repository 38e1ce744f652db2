import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import torch, bench
from shopformer_b200.synthetic import synth_windows
from shopformer_b200 import configs as CFG
cfg = sys.argv[1]
model = bench.build_model(cfg).cuda()
eng = model._sf_engine()
_, T, V = CFG.input_shape(cfg)
x = torch.from_numpy(synth_windows(65536, T, V, seed=1)[0]).cuda()
ref = eng.tokenize(x[:4096], precision="fp32")
tok = eng.tokenize(x, precision="bf16")
err = float((tok[:4096] - ref).abs().max() / ref.abs().max())
def timeit(f, n=10):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(cfg, "G", os.environ.get("SF_TOK_G"), f"tokenizer {timeit(lambda: eng.tokenize(x, precision='bf16')):.3f} ms  rel err vs fp32 {err:.2e}")
