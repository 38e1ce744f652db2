"""CPU: the tokenizer-v2 tile program (host builder in csrc/tok2_build.cu) executed by the host emulator
(csrc/tok2_emulate.cu) against the oracle's tokens.  The emulator decodes the same descriptor table the
sm_100a kernel issues and interleaves the MMA / epilogue / TMA sequences pseudo-randomly, so a missing
dependency, a deadlock or a wrong operand layout is caught here without a GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle.scoring_oracle as O
from helpers import build_model, max_abs_rel, oracle_kwargs
from shopformer_b200 import configs as CFG
from shopformer_b200 import native as N
from shopformer_b200.synthetic import synth_windows


def host_model(model):
    """sf_model with device = -1: packed on the host only (emulator fixture)."""
    lib = N.load()
    names, arrays = [], []
    for k, v in model.state_dict().items():
        if torch.is_tensor(v) and v.is_floating_point():
            names.append(k.encode())
            arrays.append(np.ascontiguousarray(v.detach().to("cpu", torch.float32).numpy()))
    n = len(names)
    h = C.c_void_p()
    cfg = model._sf_config().to_native()
    N.check(lib.sf_model_create(C.byref(cfg), n, (C.c_char_p * n)(*names), (C.c_void_p * n)(*[a.ctypes.data for a in arrays]),
                                (C.c_int64 * n)(*[a.size for a in arrays]), -1, C.byref(h)), "sf_model_create(host)")
    return lib, h


def emulate(lib, h, x, T, seed):
    fn = lib.sfdbg_tok2_emulate
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_uint32, C.c_void_p]
    S, D = C.c_int32(), C.c_int32()
    N.check(lib.sf_model_token_shape(h, T, C.byref(S), C.byref(D)), "token_shape")
    out = np.full((x.shape[0], S.value, D.value), np.nan, dtype=np.float32)
    info = (C.c_int32 * 8)()
    rc = fn(h, x.ctypes.data, x.shape[0], T, out.ctypes.data, seed, info)
    return rc, out, list(info)


@pytest.mark.parametrize("name,T,B", [("A", 24, 23), ("A", 12, 9), ("A", 18, 8), ("A", 22, 15), ("A12", 12, 15), ("P", 24, 3)])
def test_program_matches_oracle(name, T, B, dropin1, dropin2):
    model = build_model(dropin1, dropin2, name)
    kw = oracle_kwargs(model, name)
    lib, h = host_model(model)
    try:
        C_, _, V = CFG.input_shape(name)
        x = np.ascontiguousarray(synth_windows(B, T, V, seed=7)[0])
        rc, got, info = emulate(lib, h, x, T, 0)
        if name == "A12":                                  # hidden 64 (weights + activations exceed shared memory) is outside
            assert rc == -4                                # tokenizer v2: the one-window kernel serves it
            return
        N.check(rc, "sfdbg_tok2_emulate")
        # config P: adaptive average pooling of the last block's 6 time steps to 5 tokens, done by the token stage
        ref = O.tokenize(model.state_dict(), torch.from_numpy(x).double(), kw["strides"], pool_tokens=kw["pool_tokens"]).numpy()
        assert got.shape == ref.shape
        err = max_abs_rel(got, ref)
        assert err < 1e-2, (err, info)                     # bf16 operands, fp32 accumulation (GPU tests allow 2e-2)
        # any interleaving of the three sequences / eight epilogue warps gives the same bits
        for seed in (1, 2, 3):
            rc2, got2, _ = emulate(lib, h, x, T, seed)
            N.check(rc2, f"sfdbg_tok2_emulate(seed={seed})")
            assert np.array_equal(got, got2), f"schedule {seed} changed the result: a dependency is missing"
    finally:
        lib.sf_model_destroy(h)


def test_program_poisoned_window_stays_local(dropin1, dropin2):
    """A non-finite pose must only affect its own window (the mix MMA multiplies other windows' rows by zero)."""
    model = build_model(dropin1, dropin2, "A")
    lib, h = host_model(model)
    try:
        x = np.ascontiguousarray(synth_windows(10, 24, 17, seed=3)[0])
        rc, clean, _ = emulate(lib, h, x, 24, 0)
        N.check(rc, "emulate")
        x[4, 1, 7, 3] = np.inf
        rc, got, _ = emulate(lib, h, x, 24, 0)
        N.check(rc, "emulate")
        assert np.isnan(got[4]).all()
        keep = [i for i in range(10) if i != 4]
        assert np.array_equal(got[keep], clean[keep])
    finally:
        lib.sf_model_destroy(h)


def test_host_model_is_rejected_by_compute_entry_points(dropin1, dropin2):
    model = build_model(dropin1, dropin2, "A")
    lib, h = host_model(model)
    try:
        assert lib.sf_workspace_bytes(h, 16, 24) == -1
        rc = lib.sf_tokenize(h, None, 1, 24, N.SF_PREC_BF16, None, None, 0, None)
        assert rc == -1
    finally:
        lib.sf_model_destroy(h)


def test_program_dump_runs(dropin1, dropin2, capfd):
    """`sfdbg_tok2_describe` (the listing profiles/tok2_describe.py prints next to the GPU timelines) walks the same program."""
    model = build_model(dropin1, dropin2, "A")
    lib, h = host_model(model)
    try:
        lib.sfdbg_tok2_describe.restype = C.c_int
        lib.sfdbg_tok2_describe.argtypes = [C.c_void_p, C.c_int32]
        N.check(lib.sfdbg_tok2_describe(h, 24), "sfdbg_tok2_describe")
        out = capfd.readouterr().out
        assert "plan: WT=7" in out and "TOKENS" in out and "G0 " in out
        assert lib.sfdbg_tok2_describe(h, 100000) != 0          # window length outside the program's range
    finally:
        lib.sf_model_destroy(h)
