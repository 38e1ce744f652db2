"""Max relative score error of the 16-bit tensor-core path (both operand formats) and of the fp32 kernels against the
reference's float64 scores, every golden config + the reference-trained checkpoint.   python profiles/tc_error_table.py"""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200"); sys.path.insert(0, "tests")
import numpy as np, torch
from scipy.stats import spearmanr
import bench
from shopformer_b200 import configs as CFG

def rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))

rows = []
for fmt in ("fp16", "bf16"):
    os.environ["SHOPFORMER_B200_TC_FORMAT"] = fmt
    for name in ("A", "A1", "A12", "B", "C", "trained_A"):
        g = np.load(f"tests/golden/{'trained_A' if name == 'trained_A' else 'score_' + name}.npz")
        cfg = "A" if name == "trained_A" else name
        model = bench.build_model(cfg)
        if name == "trained_A":
            model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=True)
        model = model.cuda()
        eng = model._sf_engine()
        x = torch.from_numpy(g["poses"]).cuda()
        gold = g["score64"]
        T = x.shape[2]
        try:
            s_tc = eng.score_windows(x, precision="tc").cpu().numpy()
            e_tc, rho = rel(s_tc, gold), spearmanr(s_tc, gold).statistic
        except Exception as exc:
            e_tc, rho = float("nan"), float("nan")
        s_32 = eng.score_windows(x, precision="fp32").cpu().numpy()
        rows.append((fmt, name, eng.tc_formats(T), e_tc, rho, rel(s_32, gold)))
        print(f"{fmt:5s} {name:10s} formats {eng.tc_formats(T)}  tc max rel {e_tc:.2e}  spearman {rho:.6f}   fp32 kernels max rel {rel(s_32, gold):.2e}", flush=True)
