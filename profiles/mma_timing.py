import ctypes as C, sys
sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
from shopformer_b200 import native as N
lib = N.load()
out = (C.c_longlong * 2)()
for n in (16, 32, 64, 128, 256):
    for cnt in (1, 10, 100):
        for shift in (0, 17):
            lib.sfdbg_umma_timing(n, cnt, shift, out); lib.sfdbg_umma_timing(n, cnt, shift, out)
            print(f"N={n:3d} n_mma={cnt:3d} shift={shift:2d} issue={out[0]:6d} total={out[1]:6d}  per-mma={(out[1])/cnt:7.1f}")
