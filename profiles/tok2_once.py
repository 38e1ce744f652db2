"""A few launches of the tensor-core tokenizer on 65,536 config-A windows (ncu target): python profiles/tok2_once.py [n]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import torch
import bench
from shopformer_b200.synthetic import synth_windows
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = bench.build_model("A").cuda()
eng = model._sf_engine()
x = torch.from_numpy(synth_windows(65536, 24, 17, seed=1)[0]).cuda()
for _ in range(n):
    tok = eng.tokenize(x, precision="bf16")
torch.cuda.synchronize()
print("ok", float(tok.float().abs().mean()))
