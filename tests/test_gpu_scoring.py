"""GPU parity: the sm_100a scoring path (through the drop-in facades -> C ABI) against the
reference's golden vectors and against the CPU oracle on the same seeded inputs.

Tolerance (north star): floating-point scores within 1e-3 relative of the reference for the fp32
path.  The fp32 kernels are far inside that (asserted at 5e-5); window ranking / AUC are checked on
the reference-trained checkpoint where the score spread is meaningful (SURVEY finding 6).
"""
import numpy as np
import pytest
import torch

import oracle.scoring_oracle as O
from helpers import build_model, max_abs_rel, oracle_kwargs, rel_err
from shopformer_b200 import configs as CFG
from shopformer_b200 import native as N
from shopformer_b200.native import NativeError
from shopformer_b200.synthetic import synth_state_dict, synth_windows

pytestmark = pytest.mark.gpu
FP32_TOL = 5e-5          # asserted; the contract is 1e-3


@pytest.fixture(autouse=True)
def _precise_kernels_through_the_facades(monkeypatch):
    """This file checks the fp32 kernels at 5e-5 through the facades; the facades' default ("auto" = the bf16 tensor-core
    kernels) is covered by test_gpu_reference_api.py."""
    monkeypatch.setenv("SHOPFORMER_B200_PRECISION", "fp32")


def gpu_scores(model, name, x):
    with torch.no_grad():
        if CFG.variant_of(name) == 1:
            return model(x)["normality_score"]
        return model.compute_anomaly_score(x)


@pytest.mark.parametrize("name", CFG.ALL_CONFIGS)
def test_scores_match_reference_golden(name, golden_dir, dropin1, dropin2):
    g = np.load(golden_dir / f"score_{name}.npz")
    model = build_model(dropin1, dropin2, name).cuda()
    x = torch.from_numpy(g["poses"]).cuda()
    s = gpu_scores(model, name, x).cpu().numpy()
    assert s.shape == g["score64"].shape and s.dtype == np.float32
    assert rel_err(s, g["score64"]) < FP32_TOL
    n = g["tokens64"].shape[0]
    with torch.no_grad():
        tok = model.gcae.encode(x)
        rec = model.transformer(tok)
    assert max_abs_rel(tok[:n].cpu().numpy(), g["tokens64"]) < FP32_TOL
    assert max_abs_rel(rec[:n].cpu().numpy(), g["recon64"]) < FP32_TOL
    if CFG.variant_of(name) == 2:
        with torch.no_grad():
            per_tok = model.compute_anomaly_score(x, reduction="none").cpu().numpy()
        assert per_tok.shape == g["per_token64"].shape
        assert rel_err(per_tok, g["per_token64"]) < FP32_TOL
        with pytest.raises(ValueError):
            model.compute_anomaly_score(x, reduction="sum")
    else:
        with torch.no_grad():
            out = model(x, return_tokens=True)
        assert set(out.keys()) == {"normality_score", "reconstructed_tokens", "gcae_reconstructed", "tokens"}
        assert out["gcae_reconstructed"].shape == x.shape
        assert torch.equal(out["tokens"], tok)


@pytest.mark.parametrize("name,n", [("A", 1000), ("B", 1001), ("C", 517)])
def test_scores_match_oracle_ragged_batches(name, n, dropin1, dropin2):
    """Batch sizes that are not multiples of the CTA tile (16/24 windows), incl. B=1 and B=0."""
    model = build_model(dropin1, dropin2, name, seed=3)
    C, T, V = CFG.input_shape(name)
    xs, _ = synth_windows(n, T, V, seed=21)
    xs[::3] *= 1.7
    ref = O.score_windows(model.state_dict(), torch.from_numpy(xs), dtype=torch.float64, **oracle_kwargs(model, name))
    model = model.cuda()
    x = torch.from_numpy(xs).cuda()
    s = gpu_scores(model, name, x).cpu().numpy()
    assert rel_err(s, ref["score"].numpy()) < FP32_TOL
    one = gpu_scores(model, name, x[:1]).cpu().numpy()
    assert one.shape == (1,) and one[0] == s[0]          # batch composition never changes a window's score
    assert gpu_scores(model, name, x[:0]).shape == (0,)


def test_subcalls_match_oracle(dropin1, dropin2):
    """tokenize / reconstruct_tokens / compute_normality_score individually (reference API)."""
    model = build_model(dropin1, dropin2, "A")
    kw = oracle_kwargs(model, "A")
    xs, _ = synth_windows(200, 24, 17, seed=5)
    sd = {k: v.double() if v.is_floating_point() else v for k, v in model.state_dict().items()}
    tok = O.tokenize(sd, torch.from_numpy(xs).double(), kw["strides"])
    rec = O.reconstruct_v1(sd, tok, kw["nhead"])
    sc = O.score_v1(sd, tok, rec)
    model = model.cuda()
    with torch.no_grad():
        gtok = model.tokenize(torch.from_numpy(xs).cuda())
        # feed the ORACLE's tokens to isolate each stage
        grec = model.reconstruct_tokens(tok.float().cuda())
        gsc = model.compute_normality_score(tok.float().cuda(), rec.float().cuda())
    assert max_abs_rel(gtok.cpu().numpy(), tok.numpy()) < FP32_TOL
    assert max_abs_rel(grec.cpu().numpy(), rec.numpy()) < FP32_TOL
    assert rel_err(gsc.cpu().numpy(), sc.numpy()) < FP32_TOL


def test_btvc_layout_and_runtime_T(dropin1, dropin2):
    """(B,T,V,C) inputs are accepted like the reference does, and T may differ from seq_len
    (inference.py feeds 12-frame windows to 24-frame checkpoints, SURVEY finding 7)."""
    model = build_model(dropin1, dropin2, "A")
    kw = oracle_kwargs(model, "A")
    xs, _ = synth_windows(64, 12, 17, seed=9)
    cpu_sd = {k: v.clone() for k, v in model.state_dict().items()}
    ref = O.score_windows(cpu_sd, torch.from_numpy(xs), dtype=torch.float64, **kw)
    model = model.cuda()
    x = torch.from_numpy(xs).cuda()
    a = gpu_scores(model, "A", x)
    b = gpu_scores(model, "A", x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(a, b)
    assert ref["tokens"].shape[1] == 2
    assert rel_err(a.cpu().numpy(), ref["score"].numpy()) < FP32_TOL
    for T in (9, 31):
        xs, _ = synth_windows(8, T, 17, seed=T)
        ref = O.score_windows(cpu_sd, torch.from_numpy(xs), dtype=torch.float64, **kw)
        assert rel_err(gpu_scores(model, "A", torch.from_numpy(xs).cuda()).cpu().numpy(), ref["score"].numpy()) < FP32_TOL


def test_weights_are_a_derived_cache(dropin1, dropin2):
    """load_state_dict / in-place parameter updates must be picked up (BN stats move during training)."""
    model = build_model(dropin1, dropin2, "A").cuda()
    xs, _ = synth_windows(32, 24, 17, seed=2)
    x = torch.from_numpy(xs).cuda()
    s0 = gpu_scores(model, "A", x)
    eng0 = model._sf_engine()
    assert model._sf_engine() is eng0                     # unchanged weights -> same packed model
    cpu_sd = synth_state_dict(model.state_dict(), seed=5)
    model.load_state_dict(cpu_sd)
    s1 = gpu_scores(model, "A", x)
    assert model._sf_engine() is not eng0 and not torch.equal(s0, s1)
    ref = O.score_windows({k: v.cpu() for k, v in model.state_dict().items()}, torch.from_numpy(xs), dtype=torch.float64,
                          **oracle_kwargs(model, "A"))
    assert rel_err(s1.cpu().numpy(), ref["score"].numpy()) < FP32_TOL
    with torch.no_grad():
        model.gcae.encoder.layers[1].tcn.bn.running_var.mul_(1.5)
    ref = O.score_windows({k: v.cpu() for k, v in model.state_dict().items()}, torch.from_numpy(xs), dtype=torch.float64,
                          **oracle_kwargs(model, "A"))
    assert rel_err(gpu_scores(model, "A", x).cpu().numpy(), ref["score"].numpy()) < FP32_TOL


def test_trained_checkpoint_ranking_and_auc(golden_dir, dropin1):
    """Reference-trained weights: score spread is real, so ranking/AUC parity is meaningful."""
    from scipy.stats import spearmanr
    g = np.load(golden_dir / "trained_A.npz")
    model = dropin1["models"].Shopformer(**CFG.ctor_args("A"))
    sd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    s = gpu_scores(model, "A", torch.from_numpy(g["poses"]).cuda()).cpu().numpy()
    gold = g["score64"]
    assert rel_err(s, gold) < 2e-4
    rho = spearmanr(s, gold).statistic
    assert rho > 0.99999
    auc_gold = dropin1["metrics"].compute_metrics(g["labels"], gold)["auc_roc"]
    auc_ours = dropin1["metrics"].compute_metrics(g["labels"], s)["auc_roc"]
    assert abs(auc_gold - auc_ours) < 1e-6
    # and no worse than the reference's own fp32 run against its fp64 run
    assert rel_err(s, gold) < 20 * max(rel_err(g["score32"], gold), 1e-7)


def test_full_size_properties(dropin1, dropin2):
    """BASELINE config #2 size (65,536 windows): determinism, permutation equivariance and agreement
    between the fused call, the stand-alone score kernel and the host-buffer runner."""
    model = build_model(dropin1, dropin2, "A").cuda()
    n = 65536
    xs, _ = synth_windows(n, 24, 17, seed=1234)
    x = torch.from_numpy(xs).cuda()
    eng = model._sf_engine()
    s, tok, rec = eng.score_windows(x, return_tokens=True, return_recon=True)
    assert torch.equal(s, eng.score_windows(x))                                   # idempotent / deterministic
    perm = torch.randperm(n, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    assert torch.equal(eng.score_windows(x[perm]), s[perm])                       # windows are independent
    s2 = eng.normality_score(tok, rec)
    assert rel_err(s2.cpu().numpy(), s.cpu().numpy()) < 1e-5                      # stand-alone MSE kernel
    host = eng.score_host(xs, chunk=8192)
    assert np.array_equal(host, s.cpu().numpy())                                  # H2D/D2H pipeline, same kernels
    assert torch.isfinite(s).all() and float(s.min()) > 0
    sub = np.random.RandomState(0).choice(n, 512, replace=False)
    ref = O.score_windows({k: v.cpu() for k, v in model.state_dict().items()}, torch.from_numpy(xs[sub]),
                          dtype=torch.float64, **oracle_kwargs(model, "A"))
    assert rel_err(s.cpu().numpy()[sub], ref["score"].numpy()) < FP32_TOL


def test_score_windows_multi_pass(dropin1, dropin2):
    """Batches beyond 131,072 windows are scored in internal passes with a bounded workspace: same scores per window."""
    model = build_model(dropin1, dropin2, "A").cuda()
    eng = model._sf_engine()
    xs, _ = synth_windows(4096, 24, 17, seed=7)
    base = torch.from_numpy(xs).cuda()
    n = 131072 + 777
    idx = torch.arange(n, device="cuda") % 4096
    x = base[idx].contiguous()
    for prec in ("bf16", "fp32"):
        ref = eng.score_windows(base, precision=prec)
        got = eng.score_windows(x, precision=prec)
        assert got.shape == (n,)
        assert torch.equal(got, ref[idx]), prec
    assert N.load().sf_workspace_bytes(eng._h, n, 24) == N.load().sf_workspace_bytes(eng._h, 131072, 24)


def test_runner_reads_pinned_sources_in_place(dropin1, dropin2):
    """Page-locked host buffers are scored without a staging copy (the kernels read them over PCIe); pageable ones go
    through the pinned ring.  Same scores either way, equal to the device-resident call."""
    model = build_model(dropin1, dropin2, "A").cuda()
    eng = model._sf_engine()
    n = 20011
    xs, _ = synth_windows(n, 24, 17, seed=99)
    pinned = torch.from_numpy(xs).pin_memory()
    for prec in ("bf16", "fp32"):
        ref = eng.score_windows(torch.from_numpy(xs).cuda(), precision=prec).cpu().numpy()
        assert np.array_equal(eng.score_host(pinned.numpy(), precision=prec), ref), prec      # in place
        assert np.array_equal(eng.score_host(xs, precision=prec, chunk=4096), ref), prec         # staged ring


def test_errors_are_loud(dropin1, dropin2):
    model = build_model(dropin1, dropin2, "A").cuda()
    with torch.no_grad():
        with pytest.raises(ValueError):
            model(torch.zeros(2, 2, 24, 18, device="cuda"))                        # wrong keypoint count
        with pytest.raises(RuntimeError):
            model(torch.zeros(2, 2, 24, 17))                                        # CPU tensor: no CPU fallback
    eng = model._sf_engine()
    with pytest.raises(NativeError):
        eng.tokenize(torch.zeros(1, 2, 5000, 17, device="cuda"))                   # T outside the supported range


# ----------------------------------------------------------------------------------- bf16 tcgen05 path
BF16_TOL = 1e-2          # north-star tolerance for the bf16 path (relative error of the score)


@pytest.mark.parametrize("name", ["A", "A1", "A12", "B", "C"])
def test_bf16_tensor_core_path_matches_reference(name, golden_dir, dropin1, dropin2):
    g = np.load(golden_dir / f"score_{name}.npz")
    model = build_model(dropin1, dropin2, name).cuda()
    x = torch.from_numpy(g["poses"]).cuda()
    eng = model._sf_engine()
    s, tok, rec = eng.score_windows(x, precision="bf16", return_tokens=True, return_recon=True)
    n = g["tokens64"].shape[0]
    tok_err = max_abs_rel(tok[:n].cpu().numpy(), g["tokens64"])
    s_err = rel_err(s.cpu().numpy(), g["score64"])
    print(f"[bf16 {name}] tokens max|d|/max|ref| = {tok_err:.3e}, score max rel = {s_err:.3e}")
    assert tok_err < 2e-2
    assert s_err < BF16_TOL
    # fp32 path on the same inputs stays the precise one
    assert rel_err(eng.score_windows(x, precision="fp32").cpu().numpy(), g["score64"]) < FP32_TOL


def test_bf16_path_ragged_and_deterministic(dropin1, dropin2):
    model = build_model(dropin1, dropin2, "A", seed=3)
    xs, _ = synth_windows(1203, 24, 17, seed=77)
    ref = O.score_windows(model.state_dict(), torch.from_numpy(xs), dtype=torch.float64, **oracle_kwargs(model, "A"))
    model = model.cuda()
    eng = model._sf_engine()
    x = torch.from_numpy(xs).cuda()
    s = eng.score_windows(x, precision="bf16")
    assert rel_err(s.cpu().numpy(), ref["score"].numpy()) < BF16_TOL
    assert torch.equal(s, eng.score_windows(x, precision="bf16"))
    perm = torch.randperm(1203, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    assert torch.equal(eng.score_windows(x[perm], precision="bf16"), s[perm])
    assert torch.equal(eng.score_windows(x[:1], precision="bf16"), s[:1])


@pytest.mark.parametrize("name, n", [("B", 1203), ("B", 1), ("C", 257), ("A12", 515)])
def test_one_window_tokenizer_passes_are_position_independent(name, n, dropin1, dropin2):
    """Hidden-64 shapes run the one-window tokenizer, which packs several windows into a CTA pass when the shape only fits one
    CTA per SM (config B: two).  A window's score must not depend on which pass slot or which neighbour it gets: odd batch
    sizes (a last pass with one window), permutations and a batch of one give bit-identical per-window scores."""
    model = build_model(dropin1, dropin2, name, seed=5)
    _, T, V = CFG.input_shape(name)
    xs, _ = synth_windows(n, T, V, seed=78)
    ref = O.score_windows(model.state_dict(), torch.from_numpy(xs), dtype=torch.float64, **oracle_kwargs(model, name))
    model = model.cuda()
    eng = model._sf_engine()
    x = torch.from_numpy(xs).cuda()
    s = eng.score_windows(x, precision="bf16")
    assert rel_err(s.cpu().numpy(), ref["score"].numpy()) < BF16_TOL
    perm = torch.randperm(n, device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    assert torch.equal(eng.score_windows(x[perm].contiguous(), precision="bf16"), s[perm])
    assert torch.equal(eng.score_windows(x[n - 1:], precision="bf16"), s[n - 1:])
    if n > 2:
        assert torch.equal(eng.score_windows(x[1:n - 1].contiguous(), precision="bf16"), s[1:n - 1])


def test_bf16_path_trained_checkpoint_ranking(golden_dir, dropin1):
    from scipy.stats import spearmanr
    g = np.load(golden_dir / "trained_A.npz")
    model = dropin1["models"].Shopformer(**CFG.ctor_args("A"))
    model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=True)
    model = model.cuda().eval()
    s = model._sf_engine().score_windows(torch.from_numpy(g["poses"]).cuda(), precision="bf16").cpu().numpy()
    gold = g["score64"]
    err = rel_err(s, gold)
    rho = spearmanr(s, gold).statistic
    auc_gold = dropin1["metrics"].compute_metrics(g["labels"], gold)["auc_roc"]
    auc_ours = dropin1["metrics"].compute_metrics(g["labels"], s)["auc_roc"]
    print(f"[bf16 trained A] max rel {err:.3e}  spearman {rho:.6f}  auc {auc_ours:.6f} vs {auc_gold:.6f}")
    assert err < BF16_TOL and rho > 0.999 and abs(auc_gold - auc_ours) < 1e-3


def test_bf16_unsupported_shapes_are_loud(dropin1, dropin2):
    """hidden_channels = 128 (shopformer/sweep.py:24-42) is outside both tensor-core tokenizers: an explicit request for
    the tensor-core path fails loudly, the facades' "auto" policy selects the fp32 kernels."""
    model = dropin1["models"].Shopformer(**{**CFG.ctor_args("A"), "hidden_channels": 128})
    model.load_state_dict(synth_state_dict(model.state_dict(), seed=0), strict=True)
    model = model.eval().cuda()
    x = torch.from_numpy(synth_windows(4, 24, 17, seed=2)[0]).cuda()
    with pytest.raises(NativeError, match="SF_E_UNSUPPORTED"):
        model._sf_engine().score_windows(x, precision="bf16")
    auto = model._sf_engine().score_windows(x, precision="auto")
    assert torch.equal(auto, model._sf_engine().score_windows(x, precision="fp32"))


def test_adaptive_pooling_on_the_tensor_core_path(golden_dir, dropin1, dropin2):
    """Config P (shopformer_2, num_tokens = 5: AdaptiveAvgPool of the last block's 6 time steps, gcae.py:406-415) runs on
    tokenizer v2 (pooling in the token stage); its transformer (4 heads of width 34, S = 5) is outside the tensor-core
    transformer and takes the fp32 kernel -- the two kernels pick their paths independently."""
    g = np.load(golden_dir / "score_P.npz")
    model = build_model(dropin1, dropin2, "P").cuda()
    eng = model._sf_engine()
    x = torch.from_numpy(g["poses"]).cuda()
    assert eng.tc_formats(24)[0] == "f16", "config P did not reach tokenizer v2"
    s, tok, rec = eng.score_windows(x, precision="tc", return_tokens=True, return_recon=True)
    assert tuple(tok.shape[1:]) == (5, 136)
    assert max_abs_rel(tok[:g["tokens64"].shape[0]].cpu().numpy(), g["tokens64"]) < 5e-3
    assert rel_err(s.cpu().numpy(), g["score64"]) < BF16_TOL
    assert rel_err(eng.score_windows(x, precision="fp32").cpu().numpy(), g["score64"]) < FP32_TOL


@pytest.mark.parametrize("name", ["A", "A1", "A12", "B", "C"])
def test_bf16_transformer_alone_matches_oracle(name, golden_dir, dropin1, dropin2):
    """Tensor-core transformer fed with the ORACLE's tokens (isolates it from the tokenizer), incl. config C
    (12 heads, 4+4 layers, FFN 512 in four chunks, 136<->144 projections)."""
    g = np.load(golden_dir / f"score_{name}.npz")
    model = build_model(dropin1, dropin2, name)
    kw = oracle_kwargs(model, name)
    ref = O.score_windows(model.state_dict(), torch.from_numpy(g["poses"]), dtype=torch.float64, **kw)
    model = model.cuda()
    eng = model._sf_engine()
    tok = ref["tokens"].float().cuda()
    rec = eng.reconstruct_tokens(tok, precision="bf16")
    err = max_abs_rel(rec.cpu().numpy(), ref["recon"].numpy())
    s = eng.normality_score(tok, rec)
    s_err = rel_err(s.cpu().numpy(), ref["score"].numpy())
    print(f"[bf16 xf {name}] recon max|d|/max|ref| = {err:.3e}, score max rel = {s_err:.3e}")
    assert err < 2e-2 and s_err < BF16_TOL
