"""Build-container only: the reference's own train.py / evaluate.py / inference.py run UNCHANGED on top of
the drop-in packages (skipped where /root/reference is absent, e.g. on the GPU box)."""
import json
import os
import runpy
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REF = Path("/root/reference/shopformer")
PKG = Path(__file__).resolve().parent.parent / "computer-vision-shoplifting-detection_b200" / "shopformer"

pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference checkout not available")


def _run(script: str, argv, cwd, monkeypatch):
    for k in [k for k in sys.modules if k.split(".")[0] in ("models", "data", "utils")]:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.syspath_prepend(str(PKG))
    monkeypatch.setattr(sys, "argv", [script] + argv)
    monkeypatch.chdir(cwd)
    # GPU-less build container: let eval-mode calls use the training composition (test-only switch)
    monkeypatch.setenv("SHOPFORMER_B200_COMPOSITE_EVAL", "1")
    runpy.run_path(str(REF / script), run_name="__main__")
    assert "shopformer_b200" in sys.modules and str(PKG) in sys.modules["models"].__file__


def test_reference_train_evaluate_inference_run_unchanged(tmp_path, monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("host-logic test for the GPU-less container")
    torch.manual_seed(0)
    np.random.seed(0)
    out = tmp_path / "ckpt"
    _run("train.py", ["--use_synthetic", "--stage1_epochs", "1", "--stage2_epochs", "1", "--output_dir", str(out),
                      "--device", "cpu", "--batch_size", "64"], tmp_path, monkeypatch)
    assert (out / "final_model.pt").exists() and (out / "config.json").exists()
    ck = torch.load(out / "final_model.pt", map_location="cpu", weights_only=False)
    assert "gcae.encoder.layers.0.gcn.adj" in ck["model_state_dict"]
    res = tmp_path / "res.json"
    _run("inference.py", ["--checkpoint", str(out / "final_model.pt"), "--use_synthetic", "--device", "cpu",
                          "--output", str(res)], tmp_path, monkeypatch)
    r = json.load(open(res))
    assert len(r["scores"]) == 200 and 0.0 <= r["metrics"]["auc_roc"] <= 1.0
    _run("evaluate.py", ["--checkpoint", str(out / "final_model.pt"), "--use_synthetic", "--device", "cpu"], tmp_path, monkeypatch)
