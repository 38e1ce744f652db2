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
lib.sfdbg_tokenizer_timing(1, None, 0)
eng.tokenize(x, precision="bf16"); torch.cuda.synchronize()
buf = (C.c_longlong * 512)()
lib.sfdbg_tokenizer_timing(0, buf, 512)
a = np.array(buf[:480]).reshape(-1, 2)
prev = None
for i in range(len(a)):
    if a[i,0] in (135, 300, 301, 302, 138):
        print(int(a[i,0]), int(a[i,1]) - (prev or int(a[i,1]))); 
        if a[i,0] == 135: prev = int(a[i,1])
