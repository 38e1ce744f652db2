"""tcgen05.mma issue/complete timing microbenchmark (debug aid): python profiles/mma_timing.py
mode 0: no-swizzle, same A every MMA; 1: no-swizzle, rotating A tiles; 2: SWIZZLE_128B same A; 3: SWIZZLE_128B rotating A"""
import ctypes as C, sys
sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
from shopformer_b200 import native as N
lib = N.load()
out = (C.c_longlong * 2)()
for mode in (10, 11, 12):
    for n in (32, 64):
        for cnt in (16, 96):
            lib.sfdbg_umma_timing(n, cnt, 0, mode, out); lib.sfdbg_umma_timing(n, cnt, 0, mode, out)
            print(f"mode={mode} N={n:3d} n_mma={cnt:3d} issue={out[0]:6d} total={out[1]:6d}  per-mma={(out[1])/cnt:7.1f}")
