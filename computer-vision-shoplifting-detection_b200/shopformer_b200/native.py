"""ctypes binding of ``libshopformer_b200.so`` (the C ABI in ``include/shopformer_b200.h``).

This is the only place Python touches the native library.  It is deliberately thin:
plain pointers and sizes in, integer status out; torch is used by the callers for
device memory and streams only.  There is no fallback: if the library is missing or a
call fails, ``NativeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libshopformer_b200.so")

SF_OK = 0
SF_VARIANT_SHOPFORMER = 1
SF_VARIANT_SHOPFORMER_2 = 2
SF_REDUCE_MEAN = 0
SF_REDUCE_NONE = 1
SF_PREC_FP32 = 0
SF_PREC_BF16 = 1
SF_TC_AUTO, SF_TC_BF16, SF_TC_FP16 = 0, 1, 2
SF_MAX_BLOCKS = 8

ERROR_NAMES = {-1: "SF_E_INVALID", -2: "SF_E_MISSING", -3: "SF_E_SHAPE", -4: "SF_E_UNSUPPORTED",
               -5: "SF_E_CUDA", -6: "SF_E_NODEVICE"}

# every symbol include/shopformer_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "sf_abi_version", "sf_last_error", "sf_device_count", "sf_model_create", "sf_model_destroy",
    "sf_model_token_shape", "sf_workspace_bytes", "sf_tokenize", "sf_reconstruct_tokens",
    "sf_normality_score", "sf_score_windows", "sf_window_capacity", "sf_window_workspace_bytes",
    "sf_window_normalize", "sf_runner_create", "sf_runner_destroy", "sf_runner_score",
    "sf_runner_pinned_poses", "sf_selftest_umma", "sf_normalize_windows", "sf_model_tc_formats",
    "sf_score_from_tracks_workspace_bytes", "sf_score_from_tracks", "sf_runner_score_tracks",
    "sf_video_aggregate_workspace_bytes", "sf_video_aggregate", "sf_ranking_metrics_workspace_bytes", "sf_ranking_metrics",
    "sf_decoder_create", "sf_decoder_destroy", "sf_decoder_workspace_bytes", "sf_decode_poses",
)


class NativeError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, code: int, where: str, message: str):
        self.code = code
        super().__init__(f"{where}: {ERROR_NAMES.get(code, code)}: {message}")


class SfConfig(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("in_channels", C.c_int32), ("num_keypoints", C.c_int32),
        ("n_blocks", C.c_int32), ("channels", C.c_int32 * (SF_MAX_BLOCKS + 1)),
        ("strides", C.c_int32 * SF_MAX_BLOCKS), ("pool_tokens", C.c_int32), ("d_model", C.c_int32),
        ("n_heads", C.c_int32), ("n_enc_layers", C.c_int32), ("n_dec_layers", C.c_int32),
        ("d_ff", C.c_int32), ("tc_format", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class SfTracks(C.Structure):
    _fields_ = [
        ("kp_dev", C.c_void_p), ("frame_no_dev", C.c_void_p), ("track_offsets_host", C.c_void_p),
        ("track_video_host", C.c_void_p), ("gt_dev", C.c_void_p), ("gt_offsets_host", C.c_void_p),
        ("n_frames", C.c_int64), ("n_tracks", C.c_int32), ("n_videos", C.c_int32),
        ("kp_per_frame", C.c_int32), ("kp_channels", C.c_int32),
    ]


class SfWindowParams(C.Structure):
    _fields_ = [
        ("seq_len", C.c_int32), ("stride", C.c_int32), ("max_gap", C.c_int32),
        ("num_keypoints", C.c_int32), ("normalize", C.c_int32), ("add_neck", C.c_int32),
        ("include_confidence", C.c_int32), ("reserved", C.c_int32 * 1),
    ]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the library (once) and declare every prototype.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C computer-vision-shoplifting-detection_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    P = C.POINTER
    sig = {
        "sf_abi_version": (C.c_int, []),
        "sf_last_error": (C.c_char_p, []),
        "sf_device_count": (C.c_int, []),
        "sf_model_create": (C.c_int, [P(SfConfig), i32, P(C.c_char_p), P(vp), P(i64), i32, P(vp)]),
        "sf_model_destroy": (None, [vp]),
        "sf_model_token_shape": (C.c_int, [vp, i32, P(i32), P(i32)]),
        "sf_workspace_bytes": (i64, [vp, i64, i32]),
        "sf_model_tc_formats": (C.c_int, [vp, i32, P(i32), P(i32)]),
        "sf_tokenize": (C.c_int, [vp, vp, i64, i32, i32, vp, vp, i64, vp]),
        "sf_reconstruct_tokens": (C.c_int, [vp, vp, i64, i32, i32, vp, vp, i64, vp]),
        "sf_normality_score": (C.c_int, [vp, vp, vp, i64, i32, i32, vp, vp]),
        "sf_score_windows": (C.c_int, [vp, vp, i64, i32, i32, i32, vp, vp, vp, vp, i64, vp]),
        "sf_window_capacity": (i64, [P(SfTracks), P(SfWindowParams)]),
        "sf_window_workspace_bytes": (i64, [P(SfTracks), P(SfWindowParams)]),
        "sf_window_normalize": (C.c_int, [P(SfTracks), P(SfWindowParams), vp, vp, vp, vp, vp, vp, P(i64), vp, i64, vp]),
        "sf_runner_create": (C.c_int, [vp, i32, i64, P(vp)]),
        "sf_runner_destroy": (None, [vp]),
        "sf_runner_score": (C.c_int, [vp, vp, i64, i32, vp]),
        "sf_runner_pinned_poses": (vp, [vp, i32]),
        "sf_selftest_umma": (C.c_int, [i32, i32, i32, i32, vp, vp, vp]),
        "sf_normalize_windows": (C.c_int, [vp, i64, i32, i32, i32, i32, vp, vp]),
        "sf_score_from_tracks_workspace_bytes": (i64, [vp, P(SfTracks), P(SfWindowParams)]),
        "sf_score_from_tracks": (C.c_int, [vp, P(SfTracks), P(SfWindowParams), i32, vp, vp, vp, vp, P(i64), vp, i64, vp]),
        "sf_runner_score_tracks": (C.c_int, [vp, P(SfTracks), P(SfWindowParams), i32, vp, vp, vp, vp, P(i64)]),
        "sf_video_aggregate_workspace_bytes": (i64, [i64, i32]),
        "sf_video_aggregate": (C.c_int, [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, i64, vp]),
        "sf_ranking_metrics_workspace_bytes": (i64, [i64]),
        "sf_ranking_metrics": (C.c_int, [vp, vp, i64, C.c_float, P(C.c_double), vp, i64, vp]),
        "sf_decoder_create": (C.c_int, [i32, i32, i32, i32, i32, i32, P(i32), i32, P(C.c_char_p), P(vp), P(i64), i32, P(vp)]),
        "sf_decoder_destroy": (None, [vp]),
        "sf_decoder_workspace_bytes": (i64, [vp, i64, i32]),
        "sf_decode_poses": (C.c_int, [vp, vp, i64, i32, vp, vp, i64, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.sf_abi_version() != 2:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.sf_abi_version()} != 2; rebuild")
    _lib = lib
    return lib


def check(rc: int, where: str) -> None:
    if rc < 0:
        msg = load().sf_last_error()
        raise NativeError(rc, where, msg.decode() if msg else "")
