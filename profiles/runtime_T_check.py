import sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200"); sys.path.insert(0, "tests")
import numpy as np, torch
import bench
import oracle.scoring_oracle as O
from shopformer_b200.synthetic import synth_windows
model = bench.build_model("A")
sd = {k: v.clone() for k, v in model.state_dict().items()}
enc = model.gcae.encoder
model = model.cuda()
eng = model._sf_engine()
for T in (8, 16, 24, 32, 40, 48, 56):
    xs = synth_windows(300, T, 17, seed=T)[0]
    try:
        got = eng.score_windows(torch.from_numpy(xs).cuda(), precision="bf16").cpu().numpy()
    except Exception as e:
        print(T, "bf16 unsupported:", str(e)[:120]); continue
    ref = O.score_windows(sd, torch.from_numpy(xs), variant=1, strides=list(enc.strides), nhead=model.transformer.nhead, dtype=torch.float64)["score"].numpy()
    f32 = eng.score_windows(torch.from_numpy(xs).cuda(), precision="fp32").cpu().numpy()
    print(T, "bf16 max rel err", float(np.max(np.abs(got - ref) / np.abs(ref))), "fp32", float(np.max(np.abs(f32 - ref) / np.abs(ref))))
