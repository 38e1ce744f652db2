"""Precision model of the 16-bit tokenizer dataflow (torch CPU): which roundings matter for the score error."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200"); sys.path.insert(0, "tests")
import numpy as np, torch, torch.nn.functional as F
import bench
import oracle.scoring_oracle as O
torch.set_grad_enabled(False)
def q(x, kind):
    if kind == "bf16": return x.to(torch.bfloat16).to(x.dtype)
    if kind == "f16": return x.to(torch.float16).to(x.dtype)
    return x
def fold_bn(sd, p, dt):
    s = sd[p + "weight"].to(dt) / torch.sqrt(sd[p + "running_var"].to(dt) + 1e-5)
    return s, sd[p + "bias"].to(dt) - sd[p + "running_mean"].to(dt) * s
def tokenize_q(sd, poses, strides, wq, aq, adjq, dt=torch.float64, stats=None):
    prefix = "gcae.encoder."
    x = poses.to(dt)
    B, C, T, V = x.shape
    s, b = fold_bn(sd, prefix + "bn_input.", dt)
    x = x * s.view(1, C, 1, V).permute(0, 1, 2, 3).reshape(1, C, V, 1).permute(0, 1, 3, 2) + b.view(1, C, V, 1).permute(0, 1, 3, 2)
    for i, st in enumerate(strides):
        p = f"{prefix}layers.{i}."
        adj = sd[p + "gcn.adj"].to(dt); w = sd[p + "gcn.weight"].to(dt)
        if (p + "residual.0.weight") in sd:
            rs, rb = fold_bn(sd, p + "residual.1.", dt)
            rw = sd[p + "residual.0.weight"].to(dt) * rs.view(-1, 1, 1, 1)
            rbias = sd[p + "residual.0.bias"].to(dt) * rs + rb
            if i > 0: rw = q(rw, wq)
            r = F.conv2d(x, rw, rbias, stride=(st, 1))
        else:
            r = x
        if i == 0:
            g = torch.einsum("vu,bctu->bctv", adj, x)
            g = torch.einsum("bctv,co->botv", g, w) + sd[p + "gcn.bias"].to(dt).view(1, -1, 1, 1)
        else:
            a = q(adj, adjq)
            g = q(torch.einsum("vu,bctu->bctv", a, x), aq)
            g = torch.einsum("bctv,co->botv", g, q(w, wq)) + sd[p + "gcn.bias"].to(dt).view(1, -1, 1, 1)
        g = q(torch.relu(g), aq)
        ts, tb = fold_bn(sd, p + "tcn.bn.", dt)
        tw = q(sd[p + "tcn.conv.weight"].to(dt) * ts.view(-1, 1, 1, 1), wq)
        tbias = sd[p + "tcn.conv.bias"].to(dt) * ts + tb
        h = F.conv2d(g, tw, tbias, stride=(st, 1), padding=(4, 0))
        x = torch.relu(h + r)
        if stats is not None: stats.append((i, float(g.abs().max()), float(x.abs().max()), float(tw.abs().max())))
        if i + 1 < len(strides): x = q(x, aq)
    B, C, T, V = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, T, C * V)
def run(name, sd, x, gold, strides, nhead):
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    for label, wq, aq, adjq in [("exact", None, None, None), ("bf16 all (current)", "bf16", "bf16", "bf16"), ("bf16, adj exact", "bf16", "bf16", None),
                                ("w f16, act bf16, adj f16", "f16", "bf16", "f16"), ("all f16", "f16", "f16", "f16"), ("w exact, act bf16", None, "bf16", None),
                                ("w bf16, act exact", "bf16", None, None)]:
        stats = []
        tok = tokenize_q(sd, x, strides, wq, aq, adjq, stats=stats)
        rec = O.reconstruct_v1(sd64, tok, nhead)
        s = O.score_v1(sd64, tok, rec).numpy()
        rel = (s - gold) / gold
        print(f"{name:10s} {label:28s} max {np.abs(rel).max():.2e}  mean signed {rel.mean():+.2e}  rms {np.sqrt((rel**2).mean()):.2e}")
    print("   activation maxima per block (g, x, |w_tcn|):", [(i, round(a, 2), round(b, 2), round(c, 2)) for i, a, b, c in stats])
g = np.load("tests/golden/trained_A.npz")
model = bench.build_model("A")
sd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}
model.load_state_dict(sd, strict=True)
n = 512
x = torch.from_numpy(g["poses"][:n]); gold = g["score64"][:n]
run("trained_A", model.state_dict(), x, gold, list(model.gcae.encoder.strides), model.transformer.nhead)
# synthetic weights (bench config A)
from shopformer_b200.synthetic import synth_windows
m2 = bench.build_model("A")
xs = torch.from_numpy(synth_windows(512, 24, 17, seed=5)[0])
sd2 = m2.state_dict()
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd2.items()}
gold2 = O.score_windows(sd2, xs, variant=1, strides=list(m2.gcae.encoder.strides), nhead=m2.transformer.nhead, dtype=torch.float64)["score"].numpy()
run("synth_A", sd2, xs, gold2, list(m2.gcae.encoder.strides), m2.transformer.nhead)
