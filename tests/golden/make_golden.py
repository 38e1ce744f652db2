"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (it needs /root/reference):

    python tests/golden/make_golden.py            # scoring + windowing + checks
    python tests/golden/make_golden.py --train    # additionally the reference-trained checkpoint (~1 min)

What it pins:
  * score_<cfg>.npz    reference tokens / reconstruction / scores in fp64 and fp32 for the canonical
                       configs with deterministic weights (shopformer_b200.synthetic.synth_state_dict,
                       keyed by state-dict names -- no dependence on torch's RNG stream)
  * windowing.npz      the reference PoseLiftDataset classes run on a fabricated PoseLift directory
                       (shopformer_b200.synthetic.synth_poselift_video): window tensors, labels,
                       frame indices, video ids
  * trained_A.npz      a checkpoint produced by the reference's own train.py --use_synthetic, its
                       evaluation windows/labels and the reference's fp64 / fp32 scores on them
  * manifest.json      state-dict key/shape lists and seeded-init digests of the reference models
It also asserts, here, that the oracle restatement and the drop-in module trees agree with the
reference (state-dict keys, seeded init, eval outputs).
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import json
import os
import pickle
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
PKG = REPO / "computer-vision-shoplifting-detection_b200"
sys.path.insert(0, str(PKG))
sys.path.insert(0, str(REPO))

from shopformer_b200 import configs as CFG  # noqa: E402
from shopformer_b200.synthetic import synth_poselift_video, synth_state_dict, synth_windows  # noqa: E402

N_WIN = 48          # windows per config in the score goldens
N_FULL = 4          # windows whose full token / recon tensors are stored


def import_tree(root: Path, tag: str):
    """Import <root>/{models,data,utils} as private top-level packages (both reference variants use
    the same top-level names, so each gets its own sys.modules snapshot)."""
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("models", "data", "utils")}
    sys.path.insert(0, str(root))
    try:
        mods = {name: importlib.import_module(name) for name in ("models", "data.poselift_dataset", "utils.metrics")}
        if (root / "utils" / "config.py").exists():
            mods["utils.config"] = importlib.import_module("utils.config")
    finally:
        sys.path.remove(str(root))
        mine = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("models", "data", "utils")}
        sys.modules.update(saved)
    return mods, mine


def build(mods, name):
    args = CFG.ctor_args(name)
    if CFG.variant_of(name) == 1:
        return mods["models"].Shopformer(**args)
    import importlib as _il  # noqa: F401
    return mods["models"].Shopformer(args)


def digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def ref_scores(model, name, x):
    """tokens, recon, score through the reference's own entry points."""
    with torch.no_grad():
        if CFG.variant_of(name) == 1:
            out = model(x, return_tokens=True)
            return out["tokens"], out["reconstructed_tokens"], out["normality_score"], None
        _, tok = model.gcae(x)
        rec = model.transformer(tok)
        dt = x.dtype
        if dt == torch.float32:
            sc = model.compute_anomaly_score(x, "mean")
            per_tok = model.compute_anomaly_score(x, "none")
        else:
            sc = ((tok - rec) ** 2).mean(dim=(1, 2))
            per_tok = ((tok - rec) ** 2).mean(dim=2)
        return tok, rec, sc, per_tok


def make_scores(ref1, ref2, ours1, ours2, manifest):
    import oracle.scoring_oracle as O
    for name in CFG.ALL_CONFIGS:
        var = CFG.variant_of(name)
        ref, ours = (ref1, ours1) if var == 1 else (ref2, ours2)
        torch.manual_seed(0)
        rmodel = build(ref, name)
        seeded = digest(rmodel.state_dict())
        torch.manual_seed(0)
        omodel = build(ours, name)
        rsd, osd = rmodel.state_dict(), omodel.state_dict()
        assert list(rsd.keys()) == list(osd.keys()), f"{name}: state-dict keys differ"
        for k in rsd:
            assert rsd[k].shape == osd[k].shape and rsd[k].dtype == osd[k].dtype, f"{name}: {k} shape/dtype"
        assert digest(osd) == seeded, f"{name}: seeded init of the drop-in differs from the reference"
        sd = synth_state_dict(rsd, seed=0)
        rmodel.load_state_dict(sd, strict=True)
        omodel.load_state_dict(sd, strict=True)
        rmodel.eval()
        C, T, V = CFG.input_shape(name)
        xs, labels = synth_windows(N_WIN, T, V, seed=7)
        # second half: N(0,1)-scaled junk to widen the score spread and exercise ReLU sign patterns
        rs = np.random.RandomState(11)
        xs[N_WIN // 2:] = rs.randn(N_WIN - N_WIN // 2, C, T, V).astype(np.float32)
        x32 = torch.from_numpy(xs)
        tok32, rec32, sc32, pt32 = ref_scores(rmodel, name, x32)
        r64 = build(ref, name)
        r64.load_state_dict(sd, strict=True)
        r64.eval().double()
        tok64, rec64, sc64, pt64 = ref_scores(r64, name, x32.double())
        # ---- oracle restatement vs the reference (fp64 tight, fp32 loose)
        enc = rmodel.gcae.encoder
        strides = O.strides_v1(T, enc.num_tokens, len(enc.layers)) if var == 1 else O.strides_v2(T, enc.num_tokens, len(enc.layers))
        real = [l.tcn.conv.stride[0] for l in enc.layers]
        assert strides == real, f"{name}: stride rule {strides} vs reference {real}"
        pool = enc.num_tokens if (var == 2 and enc._needs_pooling) else None
        nhead = rmodel.transformer.nhead
        o64 = O.score_windows(sd, x32, variant=var, strides=strides, nhead=nhead, pool_tokens=pool, dtype=torch.float64)
        for key, refv in (("tokens", tok64), ("recon", rec64), ("score", sc64)):
            err = (o64[key] - refv).abs().max().item() / max(refv.abs().max().item(), 1e-30)
            assert err < 1e-11, f"{name}: oracle fp64 {key} rel err {err}"
        o32 = O.score_windows(sd, x32, variant=var, strides=strides, nhead=nhead, pool_tokens=pool, dtype=torch.float32)
        err32 = ((o32["score"].double() - sc64).abs() / sc64.abs()).max().item()
        ref32 = ((sc32.double() - sc64).abs() / sc64.abs()).max().item()
        assert err32 < 2e-5, f"{name}: oracle fp32 score rel err {err32}"
        if var == 2:
            on = O.score_v2(o64["tokens"], o64["recon"], "none")
            assert (on - pt64).abs().max().item() < 1e-12
        # drop-in composite path (autograd/training path) in eval mode on CPU == reference
        os.environ["SHOPFORMER_B200_COMPOSITE_EVAL"] = "1"
        omodel.eval()
        with torch.no_grad():
            if var == 1:
                osc = omodel(x32)["normality_score"]
            else:
                osc = omodel.compute_anomaly_score(x32)
        os.environ.pop("SHOPFORMER_B200_COMPOSITE_EVAL")
        cerr = ((osc.double() - sc64).abs() / sc64.abs()).max().item()
        assert cerr < 2e-5, f"{name}: drop-in composite rel err {cerr}"
        np.savez_compressed(
            HERE / f"score_{name}.npz", poses=xs, labels=labels, strides=np.asarray(strides), nhead=nhead,
            pool_tokens=0 if pool is None else pool,
            score64=sc64.numpy(), score32=sc32.numpy(),
            tokens64=tok64[:N_FULL].numpy(), recon64=rec64[:N_FULL].numpy(),
            per_token64=np.zeros(0) if pt64 is None else pt64.numpy())
        manifest["configs"][name] = {
            "variant": var, "keys": list(rsd.keys()), "shapes": [list(v.shape) for v in rsd.values()],
            "seeded_init_sha256": seeded, "params": int(sum(p.numel() for p in rmodel.parameters())),
            "ref_fp32_vs_fp64_max_rel": ref32, "oracle_fp32_vs_fp64_max_rel": err32,
            "score_mean": float(sc64.mean()), "score_std": float(sc64.std()),
        }
        print(f"[score] {name}: params={manifest['configs'][name]['params']} strides={strides} S={tok64.shape[1]} "
              f"score={float(sc64.mean()):.5f}+-{float(sc64.std()):.5f} ref32err={ref32:.2e} oracle32err={err32:.2e}")


def make_decoder(ref1, ref2, manifest):
    """GCAEDecoder.forward (shopformer/models/gcae.py:369-478; shopformer_2/models/gcae.py:425-534) of the
    reference in float64 on the reference's own tokens, every config (incl. the bilinear-resize branch where the
    transposed-conv stack overshoots seq_len)."""
    out = {}
    for name in CFG.ALL_CONFIGS:
        var = CFG.variant_of(name)
        ref = ref1 if var == 1 else ref2
        torch.manual_seed(0)
        r64 = build(ref, name)
        r64.load_state_dict(synth_state_dict(r64.state_dict(), seed=0), strict=True)
        r64.eval().double()
        C, T, V = CFG.input_shape(name)
        xs, _ = synth_windows(8, T, V, seed=21)
        xs[4:] = np.random.RandomState(5).randn(4, C, T, V).astype(np.float32)
        with torch.no_grad():
            rec, tok = r64.gcae(torch.from_numpy(xs).double())
            again = r64.gcae.decoder(tok)
        assert torch.equal(rec, again)
        out[f"{name}_poses_in"] = xs
        out[f"{name}_tokens64"] = tok.numpy()
        out[f"{name}_decoded64"] = rec.numpy()
        manifest.setdefault("decoder", {})[name] = {"tokens": list(tok.shape), "decoded": list(rec.shape),
                                                    "stack_length": int(r64.gcae.decoder.layers(torch.zeros(1, r64.gcae.decoder.initial_proj.out_features // V, tok.shape[1], V, dtype=torch.float64)).shape[2])}
        print(f"[decoder] {name}: tokens {tuple(tok.shape)} -> {tuple(rec.shape)} (conv stack length {manifest['decoder'][name]['stack_length']})")
    np.savez_compressed(HERE / "decoder.npz", **out)


def make_windowing(ref1, ref2, manifest):
    import oracle.windowing_oracle as W
    videos = {f"vid{n:02d}": synth_poselift_video(seed=100 + n, n_frames=300 + 40 * n) for n in range(2)}
    out = {}
    with tempfile.TemporaryDirectory() as td:
        root = Path(td)
        (root / "Pickle_files" / "Train").mkdir(parents=True)
        (root / "Pickle_files" / "Test").mkdir(parents=True)
        (root / "Pickle_files" / "GT").mkdir(parents=True)
        for name, (frames, gt) in videos.items():
            for split in ("Train", "Test"):
                with open(root / "Pickle_files" / split / f"{name}.pkl", "wb") as f:
                    pickle.dump(frames, f)
            np.save(root / "Pickle_files" / "GT" / f"{name}.npy", gt)
        cases = [("v1_T12", 1, dict(seq_len=12, stride=6, num_keypoints=17)),
                 ("v1_T24", 1, dict(seq_len=24, stride=12, num_keypoints=17)),
                 ("v1_T24_nonorm", 1, dict(seq_len=24, stride=12, num_keypoints=17, normalize=False)),
                 ("v1_T12_V18", 1, dict(seq_len=12, stride=6, num_keypoints=18)),
                 ("v1_T12_conf", 1, dict(seq_len=12, stride=6, num_keypoints=17, include_confidence=True)),
                 ("v2_T24", 2, dict(seq_len=24, stride=12, num_keypoints=17)),
                 ("v2_T12_neck", 2, dict(seq_len=12, stride=6, num_keypoints=18)),
                 ("v2_T12_gap9", 2, dict(seq_len=12, stride=5, num_keypoints=18, max_gap=9))]
        for tag, var, kw in cases:
            cls = (ref1 if var == 1 else ref2)["data.poselift_dataset"].PoseLiftDataset
            ds = cls(str(root), split="test", **kw)
            nch = 3 if kw.get("include_confidence") else 2
            wins = np.stack(ds.samples) if len(ds.samples) else np.zeros((0, kw["seq_len"], kw["num_keypoints"], nch), np.float32)
            out[f"{tag}_windows"] = wins.astype(np.float32)
            out[f"{tag}_labels"] = np.asarray(ds.labels, dtype=np.int64)
            item0 = ds[0][0].numpy()
            assert np.array_equal(item0, np.transpose(wins[0], (2, 0, 1)))
            if var == 2:
                out[f"{tag}_frame_indices"] = np.asarray(ds.frame_indices, dtype=np.int64)
                out[f"{tag}_video_ids"] = np.asarray(ds.video_ids)
            # oracle restatement == reference, bit for bit
            ow, ol, of = [], [], []
            for name in sorted(videos):
                frames, gt = videos[name]
                w, l, fi = W.extract_windows(frames, gt, seq_len=kw["seq_len"], stride=kw["stride"],
                                             num_keypoints=kw["num_keypoints"], variant=var,
                                             max_gap=kw.get("max_gap", 5), normalize=kw.get("normalize", True), channels=nch)
                ow += w; ol += l; of += fi
            assert ol == list(ds.labels), f"{tag}: oracle labels differ"
            assert len(ow) == len(ds.samples) and all(np.array_equal(a, b) for a, b in zip(ow, ds.samples)), f"{tag}: oracle windows differ"
            if var == 2:
                assert of == [list(x) for x in ds.frame_indices]
            else:
                out[f"{tag}_frame_indices"] = np.asarray(of, dtype=np.int64)
            manifest["windowing"][tag] = {"variant": var, **kw, "n_windows": len(ds.labels), "n_positive": int(sum(ds.labels))}
            print(f"[windowing] {tag}: {len(ds.labels)} windows, {int(sum(ds.labels))} positive")
    np.savez_compressed(HERE / "windowing.npz", **out)


def make_trained(ref1, manifest):
    """Reference train.py --use_synthetic (2+2 epochs, seeds injected) -> checkpoint + golden scores."""
    import runpy
    with tempfile.TemporaryDirectory() as td:
        argv = ["train.py", "--use_synthetic", "--stage1_epochs", "2", "--stage2_epochs", "2", "--output_dir", td,
                "--device", "cpu"]
        saved_argv, saved_mods = sys.argv, {k: sys.modules.get(k) for k in ("models", "data", "utils")}
        sys.modules.update({k: v for k, v in ref1["_mine"].items()})
        sys.argv = argv
        torch.manual_seed(0)
        np.random.seed(0)
        cwd = os.getcwd()
        os.chdir(td)
        try:
            runpy.run_path(str(REF / "shopformer" / "train.py"), run_name="__main__")
        finally:
            os.chdir(cwd)
            sys.argv = saved_argv
            for k in list(sys.modules):
                if k.split(".")[0] in ("models", "data", "utils"):
                    sys.modules.pop(k)
            sys.modules.update({k: v for k, v in saved_mods.items() if v is not None})
        ck = torch.load(Path(td) / "final_model.pt", map_location="cpu", weights_only=False)
        cfg = json.load(open(Path(td) / "config.json"))
    sd = ck["model_state_dict"]
    model = ref1["models"].Shopformer(in_channels=2, hidden_channels=cfg["hidden_channels"], latent_channels=cfg["latent_channels"],
                                      num_keypoints=17, seq_len=cfg["seq_len"], num_tokens=cfg["num_tokens"],
                                      transformer_heads=cfg["transformer_heads"], transformer_layers=cfg["transformer_layers"],
                                      transformer_ff_dim=cfg.get("transformer_ff_dim", 64), dropout=cfg["dropout"])
    model.load_state_dict(sd)
    model.eval()
    xs, labels = synth_windows(1024, cfg["seq_len"], 17, seed=99)
    x = torch.from_numpy(xs)
    with torch.no_grad():
        s32 = model(x)["normality_score"].numpy()
        s64 = model.double()(x.double())["normality_score"].numpy()
    auc = ref1["utils.metrics"].compute_metrics(labels, s64)["auc_roc"]
    flat = {f"sd::{k}": v.numpy() for k, v in sd.items()}
    np.savez_compressed(HERE / "trained_A.npz", poses=xs[:1024], labels=labels, score64=s64, score32=s32, **flat)
    manifest["trained_A"] = {"config": {k: cfg[k] for k in ("hidden_channels", "latent_channels", "seq_len", "num_tokens",
                                                            "transformer_heads", "transformer_layers", "dropout")},
                             "auc_roc_fp64": float(auc), "score_mean": float(s64.mean()), "score_std": float(s64.std()),
                             "ref_fp32_vs_fp64_max_rel": float(np.max(np.abs(s32 - s64) / np.abs(s64)))}
    print(f"[trained] A: auc={auc:.6f} score={s64.mean():.5f}+-{s64.std():.5f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.set_num_threads(8)
    ref1, mine1 = import_tree(REF / "shopformer", "ref1")
    ref1["_mine"] = mine1
    ref2, mine2 = import_tree(REF / "shopformer_2", "ref2")
    ours1, _ = import_tree(PKG / "shopformer", "ours1")
    ours2, _ = import_tree(PKG / "shopformer_2", "ours2")
    mpath = HERE / "manifest.json"
    manifest = json.load(open(mpath)) if mpath.exists() else {}
    manifest.setdefault("configs", {})
    manifest.setdefault("windowing", {})
    manifest["generator"] = {"torch": torch.__version__, "numpy": np.__version__, "reference": str(REF)}
    if a.only in ("", "scores"):
        make_scores(ref1, ref2, ours1, ours2, manifest)
    if a.only in ("", "windowing"):
        make_windowing(ref1, ref2, manifest)
    if a.only in ("", "decoder"):
        make_decoder(ref1, ref2, manifest)
    if a.train or a.only == "train":
        make_trained(ref1, manifest)
    json.dump(manifest, open(mpath, "w"), indent=1)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
