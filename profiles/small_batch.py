"""Launch-size sweep of the two bf16 kernels (setup cost, small-batch latency): python profiles/small_batch.py"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import torch
import bench
from shopformer_b200.synthetic import synth_windows
model = bench.build_model("A").cuda()
eng = model._sf_engine()
xall = torch.from_numpy(synth_windows(4096, 24, 17, seed=1)[0]).cuda()
def timeit(f, n=200):
    for _ in range(10): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n
for b in (32, 148, 296, 512, 592, 1184, 4096):
    x = xall[:b].contiguous()
    tok = eng.tokenize(x, precision="bf16")
    out = torch.empty_like(tok)
    sc = torch.empty(b, device="cuda")
    print(f"B={b:5d}  tokenizer {timeit(lambda: eng.tokenize(x, precision='bf16', out=tok)):7.1f} us   "
          f"transformer {timeit(lambda: eng.reconstruct_tokens(tok, precision='bf16', out=out)):7.1f} us   "
          f"score_windows {timeit(lambda: eng.score_windows(x, precision='bf16', out=sc)):7.1f} us")
