"""shopformer_b200 -- B200-native (sm_100a) Shopformer scoring path.

Sub-modules
    native     ctypes binding of libshopformer_b200.so (the C ABI in include/shopformer_b200.h)
    engine     packed-model owner + torch-facing calls + windowing driver
    facade     engine caching shared by the drop-in ``shopformer`` / ``shopformer_2`` facades
    modules    parameter containers / autograd (training) path shared by the drop-ins
    ingest     PoseLift pickles -> packed per-person tracks (host, one-time)
    sharding   window sharding across ranks + NCCL score all-gather
    streaming  fixed-shape per-tick scoring of concurrent camera streams (CUDA graph)
    synthetic  deterministic synthetic weights / windows / tracks for tests and bench
"""
__version__ = "0.1.0"
