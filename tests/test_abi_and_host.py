"""CPU (no GPU needed): the C-ABI library loads and exports every symbol the header declares, the
header and the ctypes mirror agree, and the host-side logic of the drop-ins behaves like the
reference's (state-dict layout, config plumbing, error behaviour, lazy output dict)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from helpers import build_model
from shopformer_b200 import configs as CFG
from shopformer_b200 import native as N
from shopformer_b200.facade import LazyOutput
from shopformer_b200.ingest import pack_videos
from shopformer_b200.synthetic import synth_poselift_video, synth_state_dict, synth_tracks, synth_windows

REPO = Path(__file__).resolve().parent.parent
HEADER = (REPO / "include" / "shopformer_b200.h").read_text()


def test_library_exports_every_declared_symbol():
    declared = set(re.findall(r"\b(sf_[a-z_0-9]+)\s*\(", HEADER))
    assert declared == set(N.ABI_SYMBOLS), declared ^ set(N.ABI_SYMBOLS)
    lib = ctypes.CDLL(N.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} not exported"
    assert N.load().sf_abi_version() == int(re.search(r"#define SF_ABI_VERSION (\d+)", HEADER).group(1))


def test_struct_mirrors_match_header():
    # sizes implied by the header: sf_config = 4 + 9 + 8 + 6 + 8 int32; sf_window_params = 8 int32
    assert ctypes.sizeof(N.SfConfig) == 4 * (4 + (N.SF_MAX_BLOCKS + 1) + N.SF_MAX_BLOCKS + 6 + 8)
    assert ctypes.sizeof(N.SfWindowParams) == 32
    assert ctypes.sizeof(N.SfTracks) == 6 * 8 + 8 + 3 * 4 + 4   # 6 pointers, int64, 3 int32 (+pad)
    assert f"#define SF_MAX_BLOCKS {N.SF_MAX_BLOCKS}" in HEADER


def test_no_device_is_a_loud_error():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = N.load()
    assert lib.sf_device_count() == -6
    assert b"no CPU fallback" in lib.sf_last_error()


@pytest.mark.parametrize("name", ["A", "B"])
def test_eval_inference_on_cpu_raises(name, dropin1, dropin2, monkeypatch):
    monkeypatch.delenv("SHOPFORMER_B200_COMPOSITE_EVAL", raising=False)
    model = build_model(dropin1, dropin2, name)
    C, T, V = CFG.input_shape(name)
    x = torch.zeros(2, C, T, V)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model(x) if name == "A" else model.compute_anomaly_score(x)


def test_training_path_is_autograd_capable(dropin1, dropin2):
    """train.py's stage-1 / stage-2 steps run through the ATen composition (SURVEY finding 8)."""
    model = dropin1["models"].Shopformer(**CFG.ctor_args("A"))
    model.train()
    x = torch.from_numpy(synth_windows(8, 24, 17, seed=0)[0])
    recon, tokens = model.gcae(x)
    loss = torch.nn.functional.mse_loss(recon, x)
    loss.backward()
    assert model.gcae.encoder.layers[0].gcn.weight.grad is not None and recon.shape == x.shape
    out = model(x)
    assert set(out.keys()) == {"normality_score", "reconstructed_tokens", "gcae_reconstructed"}
    m2 = dropin2["models"].Shopformer(CFG.ctor_args("B"))
    m2.freeze_gcae()
    m2.train()
    assert not m2.gcae.training and m2.transformer.training
    x2 = torch.from_numpy(synth_windows(4, 12, 18, seed=0)[0])
    # stage 2 (facade in train mode, frozen tokenizer in eval mode): a training step, so the encoder runs the ATen
    # composition too -- a packed native model would have to be rebuilt after every optimizer step
    loss2 = m2.compute_transformer_loss(x2)
    loss2.backward()
    assert m2.transformer.output_projection is not None and "_sf_cache" not in m2.__dict__
    assert all(p.grad is None for p in m2.gcae.parameters())
    assert any(p.grad is not None for p in m2.transformer.parameters())
    # ... while eval-mode inference on a CPU tensor still refuses: there is no CPU inference path
    m2.eval()
    with pytest.raises(RuntimeError):
        m2.compute_anomaly_score(x2)


def test_state_dict_surface(dropin1, dropin2):
    m = dropin1["models"].Shopformer(**CFG.ctor_args("A"))
    sd = m.state_dict()
    assert sd["transformer.encoder_layers.0.self_attn.in_proj_weight"].shape == (408, 136)
    assert sd["gcae.encoder.layers.3.residual.0.weight"].shape == (8, 32, 1, 1)
    assert "gcae.encoder.layers.1.residual.0.weight" in sd and sd["pos_encoder.pe"].shape == (1, 100, 136)
    m12 = dropin1["models"].Shopformer(**CFG.ctor_args("A12"))
    assert "gcae.encoder.layers.2.residual.0.weight" not in m12.state_dict()      # identity residual (stride 1, Cin=Cout)
    assert m12.gcae.encoder.strides == [2, 2, 1, 1]
    c = dropin2["models"].Shopformer(CFG.ctor_args("C"))
    assert c.state_dict()["transformer.input_projection.weight"].shape == (144, 136)
    assert c.gcae.encoder.strides == [3, 2, 2, 1] and not c.gcae.encoder._needs_pooling
    assert dropin2["models"].Shopformer(CFG.ctor_args("P")).gcae.encoder._needs_pooling
    with pytest.raises(ValueError):
        dropin1["models"].Shopformer(layout="h36m")
    assert dropin1["models"].Shopformer.from_config({"seq_len": 24, "hidden_channels": 32}).seq_len == 24


def test_lazy_output_behaves_like_a_dict():
    calls = []
    out = LazyOutput({"a": 1}, {"b": lambda: calls.append(1) or 2})
    assert out["a"] == 1 and not calls and "b" in out and len(out) == 2
    assert out["b"] == 2 and calls == [1] and out["b"] == 2 and calls == [1]
    out2 = LazyOutput({"a": 1}, {"b": lambda: 2})
    assert dict(out2) == {"a": 1, "b": 2} and sorted(out2.keys()) == ["a", "b"]
    assert LazyOutput({}, {"b": lambda: 2}).get("c", 7) == 7
    with pytest.raises(KeyError):
        LazyOutput({}, {})["x"]


def test_ingest_reproduces_reference_person_order():
    frames, gt = synth_poselift_video(seed=100)
    t = pack_videos([("v", frames, gt)])
    # persons in first-seen order 1, 7, 3, 0 ; NaN/Inf detections of person 7 (frames 60, 61) dropped
    lens = np.diff(t.track_offsets).tolist()
    assert lens == [300 - 8 - 2, 80 - 2, 15, 50]
    first = [int(t.frame_no[o]) for o in t.track_offsets[:-1]]
    assert first == [0, 50, 200, 240]
    assert all(np.all(np.diff(t.frame_no[a:b]) > 0) for a, b in zip(t.track_offsets[:-1], t.track_offsets[1:]))
    assert t.kp.shape[1:] == (17, 3) and t.kp.dtype == np.float32 and len(t.gt) == 280
    # flat (51,) and 15-keypoint detections were normalised to (17,3) rows
    i21 = int(np.where(t.frame_no[:t.track_offsets[1]] == 21)[0][0])
    i22 = int(np.where(t.frame_no[:t.track_offsets[1]] == 22)[0][0])
    assert t.kp[i21].any() and not t.kp[i22][15:].any()


def test_synthetic_generators_are_deterministic():
    a, la = synth_windows(16, 24, 17, seed=3)
    b, lb = synth_windows(16, 24, 17, seed=3)
    assert np.array_equal(a, b) and np.array_equal(la, lb) and a.shape == (16, 2, 24, 17) and a.dtype == np.float32
    assert synth_windows(4, 12, 18, seed=1)[0].shape == (4, 2, 12, 18)
    t1, t2 = synth_tracks(5, seed=2, max_len=100), synth_tracks(5, seed=2, max_len=100)
    assert all(np.array_equal(t1[k], t2[k]) for k in t1)
    m = torch.nn.Linear(4, 3)
    s1, s2 = synth_state_dict(m.state_dict(), 0), synth_state_dict(m.state_dict(), 0)
    assert all(torch.equal(s1[k], s2[k]) for k in s1) and not torch.equal(s1["weight"], synth_state_dict(m.state_dict(), 1)["weight"])


def test_torch_ops_are_registered():
    """`torch.ops.shopformer_b200.*` exist after importing the package's op shim (no compute without a GPU)."""
    import torch
    from shopformer_b200 import ops
    for name in ops.OPS:
        assert hasattr(torch.ops.shopformer_b200, name), name
    schema = str(torch.ops.shopformer_b200.score_fused.default._schema)
    assert "Tensor poses" in schema and "int model" in schema
