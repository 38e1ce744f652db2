"""GPU: pins the tcgen05 shared-memory descriptor encodings (no-swizzle K-major / MN-major operands,
16-byte row shifts) and the TMEM read-back mapping against numpy (bf16 inputs, fp32 accumulate)."""
import ctypes as C

import numpy as np
import pytest
import torch

from shopformer_b200 import native as N

pytestmark = pytest.mark.gpu


def bf16_round(a: np.ndarray) -> np.ndarray:
    return torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("mode,n,k,shift", [(0, 32, 32, 0), (0, 32, 288, 0), (0, 16, 80, 0), (0, 64, 64, 0),
                                            (0, 32, 64, 17), (0, 32, 32, 3), (1, 32, 128, 0), (1, 16, 32, 0),
                                            (1, 64, 128, 5)])
def test_umma_tile_matches_numpy(mode, n, k, shift):
    rs = np.random.RandomState(1000 * mode + n + k + shift)
    a = bf16_round(rs.randn(128 + shift, k).astype(np.float32))
    b = bf16_round(rs.randn(n, k).astype(np.float32) if mode == 0 else rs.randn(k, n).astype(np.float32))
    d = np.zeros((128, n), np.float32)
    lib = N.load()
    rc = lib.sf_selftest_umma(mode, n, k, shift, C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), C.c_void_p(d.ctypes.data))
    N.check(rc, "sf_selftest_umma")
    want = a[shift:shift + 128].astype(np.float64) @ (b.T if mode == 0 else b).astype(np.float64)
    assert np.max(np.abs(d - want)) < 1e-3 * max(1.0, np.max(np.abs(want))), (mode, n, k, shift)


def f16_round(a: np.ndarray) -> np.ndarray:
    return a.astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("mode,n,k,shift", [(12, 32, 128, 0), (12 | 1, 64, 128, 0), (12, 32, 64, 3), (12, 64, 288, 0), (12 | 1, 16, 32, 5)])
def test_umma_fp16_operands(mode, n, k, shift):
    """kind::f16 MMAs with fp16 operands (instruction-descriptor formats 0 / 0): the default format of the 16-bit
    tensor-core path.  (A mixed fp16 x bf16 descriptor is an illegal instruction on sm_100a -- measured -- which is why
    activations and weights switch format together.)"""
    rs = np.random.RandomState(77 + mode + n + k)
    mn = bool(mode & 1)
    a = f16_round(rs.randn(128 + shift, k).astype(np.float32))
    b = f16_round(rs.randn(k, n).astype(np.float32) if mn else rs.randn(n, k).astype(np.float32))
    d = np.zeros((128, n), np.float32)
    lib = N.load()
    N.check(lib.sf_selftest_umma(mode, n, k, shift, C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), C.c_void_p(d.ctypes.data)),
            "sf_selftest_umma")
    want = a[shift:shift + 128].astype(np.float64) @ (b if mn else b.T).astype(np.float64)
    assert np.max(np.abs(d - want)) < 1e-4 * max(1.0, np.max(np.abs(want))), (mode, n, k, shift)
