"""TMEM -> register read bandwidth of one SM (tcgen05.ld.32x32b), the floor of every accumulator epilogue:
python profiles/tmem_ld_bw.py   (sfdbg_tmem_ld_timing in csrc/tc_selftest.cu).
mode 0 = loads only; mode 1 = + fp16 conversion + 16-byte shared-memory stores (a conversion stage of the epilogues);
'+MMA' = while one more warp streams 128 x N x 16 tcgen05.mma with shared-memory operands (operand reads compete for
shared-memory bandwidth, accumulator writes for TMEM)."""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import torch
from shopformer_b200 import native as N
lib = N.load()
torch.zeros(1, device="cuda")
fn = lib.sfdbg_tmem_ld_timing
fn.restype = C.c_int
fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
out = (C.c_longlong * 4)()
print("mode warps  cols/ld  loads/wait  MMA N   cycles     bytes   B/cycle/SM   MMA cycles (per MMA)")
def run(mode, warps, cols, per_wait, n_mma=0, mma_n=0, n_ld=4096):
    fn(warps, 64, cols, per_wait, mode, 0, 0, out)
    rc = fn(warps, n_ld, cols, per_wait, mode, n_mma, mma_n, out)
    if rc:
        print("rc", rc); return
    mm = f"{out[3]:9d} ({out[3] / n_mma:6.1f})" if n_mma else ""
    print(f"{mode:4d} {warps:5d} {cols:8d} {per_wait:11d} {mma_n:6d} {out[0]:8d} {out[1]:9d} {out[1] / out[0]:10.1f}   {mm}")
for mode, cases in ((0, ((16, 1), (16, 2), (16, 4), (32, 1), (32, 2))), (1, ((16, 1), (16, 2), (32, 1)))):
    for warps in (1, 2, 4, 8, 16):
        for cols, per_wait in cases:
            run(mode, warps, cols, per_wait)
print("conversion (mode 1, 16 columns per load, 2 loads per wait) under a concurrent MMA stream:")
for warps in (8, 16):
    for mma_n in (64, 128, 160, 256):
        run(1, warps, 16, 2, n_mma=6000, mma_n=mma_n)
print("MMA stream alone (1 idle conversion warp doing 4 loads):")
for mma_n in (64, 128, 160, 256):
    run(0, 1, 16, 1, n_mma=6000, mma_n=mma_n, n_ld=4)
