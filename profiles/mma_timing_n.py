"""tcgen05.mma cost versus N (debug aid): straight-line issue of 128 x N x 16 MMAs, cycles from first issue to completion."""
import ctypes as C, sys
sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
from shopformer_b200 import native as N
lib = N.load()
out = (C.c_longlong * 2)()
for n in (64, 96, 128, 144, 160, 192, 224, 256):
    for cnt in (16, 96):
        lib.sfdbg_umma_timing(n, cnt, 0, 10, out); lib.sfdbg_umma_timing(n, cnt, 0, 10, out)
        print(f"N={n:3d} n_mma={cnt:3d} issue={out[0]:6d} total={out[1]:6d}  per-mma={(out[1])/cnt:7.1f}")
