"""Back-to-back timing of the two bf16 kernels on config A (65,536 windows): python profiles/kernel_times.py [config]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import torch
import bench
from shopformer_b200.synthetic import synth_windows
cfg = sys.argv[1] if len(sys.argv) > 1 else "A"
model = bench.build_model(cfg).cuda()
eng = model._sf_engine()
from shopformer_b200 import configs as CFG
_, T, V = CFG.input_shape(cfg)
x = torch.from_numpy(synth_windows(65536, T, V, seed=1)[0]).cuda()
tok = eng.tokenize(x, precision="bf16")
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(3):
    print(f"tokenizer_bf16 {timeit(lambda: eng.tokenize(x, precision='bf16')):.3f} ms   transformer_bf16 {timeit(lambda: eng.reconstruct_tokens(tok, precision='bf16')):.3f} ms   "
          f"score_windows {timeit(lambda: eng.score_windows(x, precision='bf16')):.3f} ms")
