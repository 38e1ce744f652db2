"""Prints the tokenizer-v2 tile program (groups, stages, loads) of a config: python profiles/tok2_describe.py [A] [T]
Runs on the CPU (host-only model); the indices match the stamps of profiles/tok2_timing.py."""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "computer-vision-shoplifting-detection_b200")
import bench
from test_tok2_program import host_model
name = sys.argv[1] if len(sys.argv) > 1 else "A"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 24
lib, h = host_model(bench.build_model(name))
lib.sfdbg_tok2_describe.restype = C.c_int
lib.sfdbg_tok2_describe.argtypes = [C.c_void_p, C.c_int32]
rc = lib.sfdbg_tok2_describe(h, T)
if rc:
    print("rc", rc, lib.sf_last_error().decode() if hasattr(lib, "sf_last_error") else "")
