import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import numpy as np, torch, bench
from shopformer_b200.synthetic import synth_windows
model = bench.build_model("A").cuda(); eng = model._sf_engine()
xs = torch.from_numpy(synth_windows(65536, 24, 17, seed=1)[0]).pin_memory().numpy()
for _ in range(3): eng.score_host(xs, precision="tc", chunk=8192)
res = []
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): eng.score_host(xs, precision="tc", chunk=8192)
    res.append(65536 * 10 / (time.perf_counter() - t0) / 1e6)
print("SF_RUNNER_SCHED", os.environ.get("SF_RUNNER_SCHED"), " ".join(f"{r:.2f}" for r in res), "M windows/s")
