"""GPU: tokenizer v2 (multi-window tiles, csrc/tokenizer2_bf16.cu) through the C ABI, against the oracle, against the
host emulator of its tile program (same descriptor table, so the hardware's operand decoding is pinned), and against
the one-window-per-pass kernel it replaces for the shapes it covers."""
import os

import numpy as np
import pytest
import torch

import oracle.scoring_oracle as O
from helpers import build_model, max_abs_rel, oracle_kwargs
from shopformer_b200 import native as N
from shopformer_b200.synthetic import synth_windows
from test_tok2_program import emulate, host_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model_a(dropin1, dropin2):
    return build_model(dropin1, dropin2, "A")           # stays on the CPU (oracle / emulator side)


@pytest.fixture(scope="module")
def eng_a(dropin1, dropin2):
    m = build_model(dropin1, dropin2, "A").cuda()
    eng = m._sf_engine()
    eng._keepalive = m
    return eng


def test_tok2_is_the_kernel_that_runs(eng_a):
    lib = N.load()
    eng = eng_a
    import ctypes as C
    before, after = (C.c_longlong * 5)(), (C.c_longlong * 5)()
    lib.sfdbg_launch_counts(before, 5)                  # host-side launch counters: [tok2, one-window tc, fp32, ...]
    x = torch.from_numpy(synth_windows(64, 24, 17, seed=1)[0]).cuda()
    eng.tokenize(x, precision="bf16")
    torch.cuda.synchronize()
    lib.sfdbg_launch_counts(after, 5)
    assert after[0] == before[0] + 1 and after[1] == before[1] and after[2] == before[2], "tokenizer v2 did not run for config A"


@pytest.mark.parametrize("T,B", [(24, 23), (24, 1), (24, 7), (24, 8), (12, 50), (18, 9), (22, 15)])
def test_tok2_matches_oracle_and_emulator(model_a, eng_a, T, B):
    kw = oracle_kwargs(model_a, "A")
    xs = np.ascontiguousarray(synth_windows(B, T, 17, seed=11)[0])
    ref = O.tokenize(model_a.state_dict(), torch.from_numpy(xs).double(), kw["strides"]).numpy()
    lib, h = host_model(model_a)
    try:
        rc, emu, _ = emulate(lib, h, xs, T, 0)
        N.check(rc, "emulate")
    finally:
        lib.sf_model_destroy(h)
    eng = eng_a
    got = eng.tokenize(torch.from_numpy(xs).cuda(), precision="bf16").cpu().numpy()
    assert max_abs_rel(got, ref) < 1e-2
    # same program, fp32 accumulation in a different order: an intermediate may round to the neighbouring bf16 (one ulp =
    # 0.4 % of that element), which is all that may differ
    assert max_abs_rel(got, emu) < 4e-3, "hardware disagrees with the emulated tile program"


def test_tok2_large_ragged_deterministic_and_position_independent(model_a, eng_a):
    kw = oracle_kwargs(model_a, "A")
    xs = synth_windows(1203, 24, 17, seed=77)[0]
    eng = eng_a
    x = torch.from_numpy(xs).cuda()
    t = eng.tokenize(x, precision="bf16")
    ref = O.tokenize(model_a.state_dict(), torch.from_numpy(xs).double(), kw["strides"]).numpy()
    assert max_abs_rel(t.cpu().numpy(), ref) < 1e-2
    assert torch.equal(t, eng.tokenize(x, precision="bf16"))
    perm = torch.randperm(1203, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    assert torch.equal(eng.tokenize(x[perm].contiguous(), precision="bf16"), t[perm])
    assert torch.equal(eng.tokenize(x[5:6].contiguous(), precision="bf16"), t[5:6])


def test_tok2_many_tiles_per_cta(eng_a):
    """More tiles than SMs x 2: every CTA runs several tiles (barrier phases, pose prefetch, bulk-store drain)."""
    xs = synth_windows(7 * 148 * 3 + 5, 24, 17, seed=5)[0]
    eng = eng_a
    x = torch.from_numpy(xs).cuda()
    t = eng.tokenize(x, precision="bf16")
    # the same windows in one-tile launches
    idx = [0, 7 * 148, 7 * 148 * 2 + 3, xs.shape[0] - 1]
    for i in idx:
        assert torch.equal(eng.tokenize(x[i:i + 1].contiguous(), precision="bf16"), t[i:i + 1])


def test_tok2_non_finite_pose_stays_in_its_window(eng_a):
    xs = synth_windows(20, 24, 17, seed=3)[0]
    eng = eng_a
    clean = eng.tokenize(torch.from_numpy(xs).cuda(), precision="bf16").cpu().numpy()
    xs[4, 1, 7, 3] = np.inf
    xs[12, 0, 0, 0] = np.nan
    got = eng.tokenize(torch.from_numpy(xs).cuda(), precision="bf16").cpu().numpy()
    assert np.isnan(got[4]).all() and np.isnan(got[12]).all()
    keep = [i for i in range(20) if i not in (4, 12)]
    assert np.array_equal(got[keep], clean[keep])


def test_one_window_kernel_still_serves_the_same_shape(model_a, eng_a):
    """SF_TOK2_OFF routes config A to tokenizer_bf16_kernel (what hidden-64 shapes use): both meet the tolerance."""
    kw = oracle_kwargs(model_a, "A")
    xs = synth_windows(300, 24, 17, seed=9)[0]
    ref = O.tokenize(model_a.state_dict(), torch.from_numpy(xs).double(), kw["strides"]).numpy()
    eng = eng_a
    x = torch.from_numpy(xs).cuda()
    new = eng.tokenize(x, precision="bf16").cpu().numpy()
    os.environ["SF_TOK2_OFF"] = "1"
    try:
        old = eng.tokenize(x, precision="bf16").cpu().numpy()
    finally:
        del os.environ["SF_TOK2_OFF"]
    assert max_abs_rel(old, ref) < 1e-2 and max_abs_rel(new, ref) < 1e-2
