"""GPU: the path a reference user reaches -- `model(poses)`, `compute_anomaly_score`, and the reference's own
train.py / evaluate.py / inference.py -- runs on the native sm_100a kernels, by default on the bf16 tcgen05 ones.

The reference scripts are staged UNMODIFIED into tests/_ref_scripts/ by `__graft_entry__.build()` (git-ignored; the
directory travels with the gpurun snapshot).  Where they are absent the loops are replayed verbatim instead:
shopformer/evaluate.py:83-104, shopformer/inference.py:67-94, shopformer/train.py:300-331,
shopformer_2/train.py:237-263, shopformer_2/evaluate.py:36-118."""
import ctypes as C
import json
import os
import runpy
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle.scoring_oracle as O
from helpers import build_model, oracle_kwargs, rel_err
from shopformer_b200 import configs as CFG
from shopformer_b200 import native as N
from shopformer_b200.synthetic import synth_windows

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent
PKG = REPO / "computer-vision-shoplifting-detection_b200"
STAGED = REPO / "tests" / "_ref_scripts"


def _launches():
    """Host-side launch counters of the library: [tokenizer v2, one-window tc tokenizer, fp32 tokenizer, tc transformer, fp32 transformer]."""
    lib = N.load()
    buf = (C.c_longlong * 5)()
    lib.sfdbg_launch_counts(buf, 5)
    return list(buf)


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_default_precision_is_the_tensor_core_path(name, golden_dir, dropin1, dropin2, monkeypatch):
    monkeypatch.delenv("SHOPFORMER_B200_PRECISION", raising=False)
    g = np.load(golden_dir / f"score_{name}.npz")
    model = build_model(dropin1, dropin2, name).cuda()
    x = torch.from_numpy(g["poses"]).cuda()
    v1 = CFG.variant_of(name) == 1

    def facade_scores():
        with torch.no_grad():
            return model(x)["normality_score"] if v1 else model.compute_anomaly_score(x)

    s_auto = facade_scores()
    eng = model._sf_engine()
    assert torch.equal(s_auto, eng.score_windows(x, precision="bf16")), "the facade default is not the bf16 kernels"
    assert rel_err(s_auto.cpu().numpy(), g["score64"]) < 1e-2
    # sub-module entry points follow the same policy
    with torch.no_grad():
        tok = model.gcae.encode(x)
        rec = model.transformer(tok)
    assert torch.equal(tok, eng.tokenize(x, precision="bf16"))
    assert torch.equal(rec, eng.reconstruct_tokens(tok, precision="bf16"))
    # attribute and environment select the precise kernels
    model.sf_precision = "fp32"
    s32 = facade_scores()
    assert rel_err(s32.cpu().numpy(), g["score64"]) < 5e-5
    model.sf_precision = None
    monkeypatch.setenv("SHOPFORMER_B200_PRECISION", "fp32")
    assert torch.equal(facade_scores(), s32)
    monkeypatch.setenv("SHOPFORMER_B200_PRECISION", "fp16")
    with pytest.raises(ValueError):
        facade_scores()


def test_auto_falls_back_to_fp32_kernels_for_uncovered_shapes(dropin1, dropin2, monkeypatch):
    monkeypatch.delenv("SHOPFORMER_B200_PRECISION", raising=False)
    model = build_model(dropin1, dropin2, "P")             # adaptive pooling: outside the tensor-core kernels
    kw = oracle_kwargs(model, "P")
    xs = synth_windows(40, 24, 17, seed=4)[0]
    ref = O.score_windows(model.state_dict(), torch.from_numpy(xs), dtype=torch.float64, **kw)["score"].numpy()
    model = model.cuda()
    s = model.compute_anomaly_score(torch.from_numpy(xs).cuda()).cpu().numpy()
    assert rel_err(s, ref) < 5e-5


def test_frozen_tokenizer_in_training_mode_does_not_rebuild_the_engine(dropin2, monkeypatch):
    """shopformer_2 stage 2: facade in train mode, GCAE frozen in eval mode (train.py:266-429).  The encoder then runs the
    ATen composition: a native engine would be re-packed after every optimizer step."""
    monkeypatch.delenv("SHOPFORMER_B200_PRECISION", raising=False)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = dropin2["models"].Shopformer(CFG.ctor_args("B")).cuda()
    model.freeze_gcae()
    model.train()
    x = torch.from_numpy(synth_windows(8, 12, 18, seed=1)[0]).cuda()
    loss = model.compute_transformer_loss(x)
    loss.backward()
    assert "_sf_cache" not in model.__dict__


def _run_script(path: Path, argv, cwd, monkeypatch, pkg):
    for k in [k for k in sys.modules if k.split(".")[0] in ("models", "data", "utils")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.syspath_prepend(str(PKG / pkg))
    monkeypatch.setattr(sys, "argv", [path.name] + argv)
    monkeypatch.chdir(cwd)
    monkeypatch.delenv("SHOPFORMER_B200_COMPOSITE_EVAL", raising=False)
    runpy.run_path(str(path), run_name="__main__")
    assert str(PKG / pkg) in sys.modules["models"].__file__


def test_reference_scripts_run_unchanged_on_the_native_kernels(tmp_path, monkeypatch, dropin1):
    monkeypatch.delenv("SHOPFORMER_B200_PRECISION", raising=False)
    lib = N.load()
    torch.manual_seed(0)
    np.random.seed(0)
    staged = STAGED / "shopformer"
    out = tmp_path / "ckpt"
    if (staged / "train.py").exists():
        before = _launches()
        _run_script(staged / "train.py", ["--use_synthetic", "--stage1_epochs", "1", "--stage2_epochs", "1", "--output_dir", str(out),
                                          "--batch_size", "64"], tmp_path, monkeypatch, "shopformer")
        assert (out / "final_model.pt").exists() and (out / "config.json").exists()
        after = _launches()
        assert after[0] > before[0] and after[3] > before[3], "train.py's evaluate() did not reach the tensor-core kernels"
        res = tmp_path / "res.json"
        _run_script(staged / "inference.py", ["--checkpoint", str(out / "final_model.pt"), "--use_synthetic", "--output", str(res)],
                    tmp_path, monkeypatch, "shopformer")
        r = json.load(open(res))
        assert len(r["scores"]) == 200 and 0.0 <= r["metrics"]["auc_roc"] <= 1.0 and np.isfinite(r["scores"]).all()
        _run_script(staged / "evaluate.py", ["--checkpoint", str(out / "final_model.pt"), "--use_synthetic", "--output",
                                             str(tmp_path / "training_results.json"), "--include_scores"], tmp_path, monkeypatch, "shopformer")
        ev = json.load(open(tmp_path / "training_results.json"))
        assert "auc_roc" in json.dumps(ev)
        ckpt = torch.load(out / "final_model.pt", map_location="cpu", weights_only=False)
        cfg = json.load(open(out / "config.json"))
        model = dropin1["models"].Shopformer(**{k: cfg[k] for k in ("hidden_channels", "latent_channels", "num_keypoints", "seq_len", "num_tokens",
                                                                 "transformer_heads", "transformer_layers", "dropout")})
        model.load_state_dict(ckpt["model_state_dict"])
    else:
        model = build_model(dropin1, None, "A")
    # the three loops, verbatim, on the checkpoint: native scores == fp64 oracle within the bf16 contract
    model = model.cuda().eval()
    ds = dropin1["data"].SyntheticPoseLiftDataset(num_samples=96, seq_len=model.seq_len, anomaly_ratio=0.3)
    scores_eval = []
    with torch.no_grad():
        for i in range(len(ds)):                                            # evaluate.py:89-99 (batch = 1)
            poses, label = ds[i]
            poses = poses.unsqueeze(0).to("cuda")
            scores_eval.append(model(poses)["normality_score"].cpu().numpy()[0])
    loader = torch.utils.data.DataLoader(ds, batch_size=32, shuffle=False)
    scores_train = []
    with torch.no_grad():
        for poses, labels in loader:                                        # train.py:311-319 (batch = 32)
            scores_train.extend(model(poses.to("cuda"))["normality_score"].cpu().numpy())
    x = torch.stack([ds[i][0] for i in range(len(ds))])
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    enc = model.gcae.encoder
    gold = O.score_windows(sd, x, variant=1, strides=list(enc.strides), nhead=model.transformer.nhead, dtype=torch.float64)["score"].numpy()
    assert rel_err(np.asarray(scores_eval), gold) < 1e-2
    assert np.array_equal(np.asarray(scores_eval), np.asarray(scores_train)), "a window's score must not depend on the batch it is in"


def test_shopformer2_evaluation_loops_on_the_native_kernels(dropin2, monkeypatch):
    """evaluate_frame_level / evaluate_video_level of shopformer_2/evaluate.py:36-118 and train.py:237-263, replayed on
    the drop-in dataset + model (the script itself imports matplotlib, which this image does not have)."""
    monkeypatch.delenv("SHOPFORMER_B200_PRECISION", raising=False)
    model = build_model(None, dropin2, "B").cuda()
    xs, labels = synth_windows(100, 12, 18, seed=8)
    ds = torch.utils.data.TensorDataset(torch.from_numpy(xs), torch.from_numpy(labels.astype(np.int64)))
    loader = torch.utils.data.DataLoader(ds, batch_size=32, shuffle=False, num_workers=0)
    all_scores, all_labels = [], []
    model.eval()
    with torch.no_grad():
        for poses, lab in loader:
            poses = poses.to("cuda")
            scores = model.compute_anomaly_score(poses)
            all_scores.extend(scores.cpu().numpy())
            all_labels.extend(lab.numpy())
    all_scores = np.array(all_scores)
    gold = O.score_windows({k: v.cpu() for k, v in model.state_dict().items()}, torch.from_numpy(xs), dtype=torch.float64,
                           **oracle_kwargs(model, "B"))["score"].numpy()
    assert rel_err(all_scores, gold) < 1e-2
    m = dropin2["metrics"].compute_metrics(np.array(all_labels), all_scores)
    assert 0.0 <= m["auc_roc"] <= 1.0
    video_scores = defaultdict(list)
    for i, s in enumerate(all_scores):
        video_scores[f"v{i // 10}"].append(float(s))
    video_labels = {k: int(i % 2) for i, k in enumerate(video_scores)}
    vm = dropin2["metrics"].compute_video_level_metrics(video_scores, video_labels, "max")
    assert "auc_roc" in vm


def test_stage2_training_keeps_the_frozen_tokenizer_on_the_native_kernels(dropin2):
    """shopformer_2 stage 2 (shopformer_2/train.py:266-429, models/shopformer.py:94-101): GCAE frozen and in eval mode, the
    transformer trains.  `encode` must reach the native tokenizer (first item) through a packed model keyed
    on the encoder alone -- optimizer steps on the transformer must not rebuild it -- and give the eval-mode tokens."""
    import oracle.scoring_oracle as O
    from helpers import build_model, max_abs_rel, oracle_kwargs
    model = build_model(None, dropin2, "B").cuda()
    model.sf_precision = "fp32"
    x = torch.from_numpy(synth_windows(64, 12, 18, seed=4)[0]).cuda()
    ref = O.tokenize({k: v.cpu() for k, v in model.state_dict().items()}, x.cpu().double(), oracle_kwargs(model, "B")["strides"]).numpy()
    model.freeze_gcae()
    model.train()
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-3)
    engines = set()
    for _ in range(3):
        tokens = model.encode(x)                                   # no_grad inside, frozen tokenizer
        assert not tokens.requires_grad
        assert max_abs_rel(tokens.cpu().numpy(), ref) < 5e-5       # eval-mode BatchNorm statistics, native fp32 kernels
        engines.add(id(model.__dict__["_sf_tok_cache"][1]))
        loss = ((model.transformer(tokens) - tokens) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
    assert len(engines) == 1, "the tokenizer's packed model was rebuilt by transformer updates"
    assert "_sf_cache" not in model.__dict__, "the full packed model must not be built while training"
