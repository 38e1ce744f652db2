"""Host-side owner of a packed native model (``sf_model``) + the torch-facing calls.

``ScoringEngine`` is what the drop-in facades hold: it turns a reference-format
``state_dict`` + config into a ``sf_model`` (BatchNorm folding happens in native code),
and exposes the C-ABI calls on torch CUDA tensors (torch supplies memory and the current
stream, nothing else).  All compute is in ``libshopformer_b200.so``; a failure of the
native library is an exception, never a fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import native as N


@dataclass
class EngineConfig:
    """Everything sf_config needs (mirrors the reference constructors' arguments)."""
    variant: int
    in_channels: int
    num_keypoints: int
    channels: List[int]            # [in, hidden, ..., latent]
    strides: List[int]
    d_model: int
    n_heads: int
    n_enc_layers: int
    n_dec_layers: int
    d_ff: int
    pool_tokens: int = 0
    # 16-bit operand format of the tensor-core path: "auto" (fp16 when every packed weight is inside fp16's range, else
    # bf16) | "bf16" (wide range, 8 significant bits) | "fp16".  Env SHOPFORMER_B200_TC_FORMAT overrides "auto".
    tc_format: str = "auto"

    def to_native(self) -> N.SfConfig:
        c = N.SfConfig()
        c.variant, c.in_channels, c.num_keypoints = self.variant, self.in_channels, self.num_keypoints
        c.n_blocks = len(self.strides)
        if c.n_blocks > N.SF_MAX_BLOCKS or len(self.channels) != c.n_blocks + 1:
            raise ValueError("bad tokenizer depth / channel list")
        for i, v in enumerate(self.channels):
            c.channels[i] = int(v)
        for i, v in enumerate(self.strides):
            c.strides[i] = int(v)
        c.pool_tokens, c.d_model, c.n_heads = self.pool_tokens, self.d_model, self.n_heads
        c.n_enc_layers, c.n_dec_layers, c.d_ff = self.n_enc_layers, self.n_dec_layers, self.d_ff
        fmt = self.tc_format
        if fmt == "auto":
            import os
            fmt = os.environ.get("SHOPFORMER_B200_TC_FORMAT", "auto").lower()
        try:
            c.tc_format = {"auto": N.SF_TC_AUTO, "bf16": N.SF_TC_BF16, "fp16": N.SF_TC_FP16, "f16": N.SF_TC_FP16}[fmt]
        except KeyError:
            raise ValueError(f"unknown tensor-core operand format {fmt!r} (auto | bf16 | fp16)") from None
        return c


def _stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class ScoringEngine:
    """One immutable packed model on one GPU."""

    def __init__(self, cfg: EngineConfig, state_dict: Dict[str, torch.Tensor], device: torch.device):
        lib = N.load()
        N.check(lib.sf_device_count(), "sf_device_count")
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ScoringEngine needs a CUDA device; there is no CPU path")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        names, arrays = [], []
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or not v.is_floating_point():
                continue
            names.append(k.encode())
            arrays.append(np.ascontiguousarray(v.detach().to("cpu", torch.float32).numpy()))
        n = len(names)
        c_names = (C.c_char_p * n)(*names)
        c_ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
        c_numel = (C.c_int64 * n)(*[a.size for a in arrays])
        handle = C.c_void_p()
        ncfg = cfg.to_native()
        N.check(lib.sf_model_create(C.byref(ncfg), n, c_names, c_ptrs, c_numel, idx, C.byref(handle)), "sf_model_create")
        self._lib = lib
        self._h = handle
        self._runners: Dict[Tuple[int, int], C.c_void_p] = {}
        self._ws: Dict[int, torch.Tensor] = {}              # stream -> reusable workspace (grow-only)
        self._auto: Dict[Tuple[str, int], str] = {}         # ("tok" | "xf", T or S) -> precision picked by "auto"
        self._shape_cache: Dict[int, Tuple[int, int]] = {}

    # -- lifetime
    def close(self) -> None:
        if getattr(self, "_h", None):
            for r in self._runners.values():
                self._lib.sf_runner_destroy(r)
            self._runners.clear()
            self._lib.sf_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- shapes
    def token_shape(self, T: int) -> Tuple[int, int]:
        hit = self._shape_cache.get(T)
        if hit is None:
            s, d = C.c_int32(), C.c_int32()
            N.check(self._lib.sf_model_token_shape(self._h, T, C.byref(s), C.byref(d)), "sf_model_token_shape")
            hit = self._shape_cache[T] = (s.value, d.value)
        return hit

    def _workspace(self, B: int, T: int) -> Tuple[Optional[torch.Tensor], int]:
        """Reusable per-stream workspace (grow-only): kernels of consecutive calls on one stream are ordered, so
        the buffer of the previous call is free by the time the next call's kernels read it."""
        nbytes = self._lib.sf_workspace_bytes(self._h, B, T)
        N.check(int(min(nbytes, 0)), "sf_workspace_bytes")
        if nbytes == 0:
            return None, 0
        key = torch.cuda.current_stream(self.device).cuda_stream
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = self._ws[key] = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return buf, int(nbytes)

    def tc_formats(self, T: int) -> Tuple[str, str]:
        """Operand formats ("f16" | "bf16") of the tensor-core tokenizer / transformer for windows of T frames."""
        a, b = C.c_int32(), C.c_int32()
        N.check(self._lib.sf_model_tc_formats(self._h, T, C.byref(a), C.byref(b)), "sf_model_tc_formats")
        return ("f16" if a.value else "bf16", "f16" if b.value else "bf16")

    # -- precision policy: "fp32" | "tc" (alias "bf16": the 16-bit tensor-core kernels; their operand format is the
    # model's tc_format) | "auto" (tensor-core kernels when they cover the shape, else fp32)
    def _run_prec(self, kind: str, key: int, precision: str, call):
        if precision in self._PREC:
            return call(self._PREC[precision])
        if precision != "auto":
            raise ValueError(f"unknown precision {precision!r} (fp32 | tc | bf16 | auto)")
        pick = self._auto.get((kind, key))
        if pick is None:
            try:
                out = call(N.SF_PREC_BF16)
                self._auto[(kind, key)] = "bf16"
                return out
            except N.NativeError as exc:
                if exc.code != -4:                 # only SF_E_UNSUPPORTED selects the other kernels
                    raise
                self._auto[(kind, key)] = pick = "fp32"
        return call(self._PREC[pick])

    def _poses(self, poses: torch.Tensor) -> torch.Tensor:
        if poses.dim() != 4:
            raise ValueError(f"poses must be (B,C,T,V), got {tuple(poses.shape)}")
        if poses.shape[1] != self.cfg.in_channels or poses.shape[3] != self.cfg.num_keypoints:
            raise ValueError(f"poses must be (B,{self.cfg.in_channels},T,{self.cfg.num_keypoints}), got {tuple(poses.shape)}")
        if not poses.is_cuda:
            raise RuntimeError("native scoring path takes CUDA tensors only (no CPU fallback)")
        return poses.to(self.device, torch.float32).contiguous()

    # -- the four reference-facing calls
    _PREC = {"fp32": N.SF_PREC_FP32, "bf16": N.SF_PREC_BF16, "tc": N.SF_PREC_BF16, "tc16": N.SF_PREC_BF16}

    def tokenize(self, poses: torch.Tensor, precision: str = "fp32", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self._poses(poses)
        B, _, T, _ = x.shape
        S, D = self.token_shape(T)
        if out is None:
            out = torch.empty(B, S, D, dtype=torch.float32, device=self.device)
        elif out.shape != (B, S, D) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("`out` must be a contiguous fp32 (B,S,D) tensor")
        ws, nb = self._workspace(B, T)
        st = _stream_ptr(self.device)
        self._run_prec("tok", T, precision, lambda p: N.check(
            self._lib.sf_tokenize(self._h, _ptr(x), B, T, p, _ptr(out), _ptr(ws), nb, st), "sf_tokenize"))
        return out

    def reconstruct_tokens(self, tokens: torch.Tensor, precision: str = "fp32", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        t = self._tokens(tokens, "tokens")
        B, S, D = t.shape
        if out is None:
            out = torch.empty_like(t)
        elif out.shape != t.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.device:
            raise ValueError("`out` must be a contiguous fp32 (B,S,D) tensor on the engine's device")
        st = _stream_ptr(self.device)
        self._run_prec("xf", S, precision, lambda p: N.check(
            self._lib.sf_reconstruct_tokens(self._h, _ptr(t), B, S, p, _ptr(out), None, 0, st), "sf_reconstruct_tokens"))
        return out

    def _tokens(self, t: torch.Tensor, what: str) -> torch.Tensor:
        """(B,S,D) fp32 on the engine's device with the model's token width (the C ABI only receives S)."""
        if t.dim() != 3 or t.shape[2] != self.cfg.channels[-1] * self.cfg.num_keypoints:
            raise ValueError(f"{what} must be (B,S,{self.cfg.channels[-1] * self.cfg.num_keypoints}), got {tuple(t.shape)}")
        if not 1 <= t.shape[1] <= 100:
            raise ValueError(f"{what}: S={t.shape[1]} outside [1,100] (positional-encoding table)")
        if not t.is_cuda:
            raise RuntimeError("native scoring path takes CUDA tensors only (no CPU fallback)")
        return t.to(self.device, torch.float32).contiguous()

    def normality_score(self, tokens: torch.Tensor, recon: torch.Tensor, reduction: str = "mean") -> torch.Tensor:
        t = self._tokens(tokens, "tokens")
        if recon.shape != tokens.shape:
            raise ValueError(f"recon {tuple(recon.shape)} must have the shape of tokens {tuple(tokens.shape)}")
        r = self._tokens(recon, "recon")
        B, S, _ = t.shape
        red = self._reduction(reduction)
        out = torch.empty((B,) if red == N.SF_REDUCE_MEAN else (B, S), dtype=torch.float32, device=self.device)
        N.check(self._lib.sf_normality_score(self._h, _ptr(t), _ptr(r), B, S, red, _ptr(out), _stream_ptr(self.device)),
                "sf_normality_score")
        return out

    @staticmethod
    def _reduction(reduction: str) -> int:
        if reduction == "mean":
            return N.SF_REDUCE_MEAN
        if reduction == "none":
            return N.SF_REDUCE_NONE
        raise ValueError(f"Unknown reduction: {reduction}")

    def score_windows(self, poses: torch.Tensor, reduction: str = "mean", precision: str = "fp32",
                      return_tokens: bool = False, return_recon: bool = False, out: Optional[torch.Tensor] = None):
        """poses -> scores (and optionally tokens / reconstruction) in one native call.  ``out`` lets the
        caller have the scores written straight into a slice of a larger buffer (e.g. the all-gather buffer)."""
        x = self._poses(poses)
        B, _, T, _ = x.shape
        S, D = self.token_shape(T)
        red = self._reduction(reduction)
        shape = (B,) if red == N.SF_REDUCE_MEAN else (B, S)
        if out is not None:
            if out.shape != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.device:
                raise ValueError("`out` must be a contiguous fp32 tensor of the score shape on the engine's device")
            scores = out
        else:
            scores = torch.empty(shape, dtype=torch.float32, device=self.device)
        tok = torch.empty(B, S, D, dtype=torch.float32, device=self.device) if return_tokens else None
        rec = torch.empty(B, S, D, dtype=torch.float32, device=self.device) if return_recon else None
        ws, nb = self._workspace(B, T)
        st = _stream_ptr(self.device)
        self._run_prec("score", T, precision, lambda p: N.check(
            self._lib.sf_score_windows(self._h, _ptr(x), B, T, red, p, _ptr(scores), _ptr(tok), _ptr(rec), _ptr(ws), nb, st),
            "sf_score_windows"))
        if return_tokens or return_recon:
            return scores, tok, rec
        return scores

    # -- host-buffer path (what a reference-side loop calls per batch)
    def score_host(self, poses: np.ndarray, precision: str = "fp32", chunk: int = 16384) -> np.ndarray:
        """numpy (B,C,T,V) fp32 on the host -> numpy scores (B,); H2D, kernels and D2H are
        pipelined in chunks inside the native runner (two streams, pinned staging)."""
        a = np.ascontiguousarray(poses, dtype=np.float32)
        if a.ndim != 4 or a.shape[1] != self.cfg.in_channels or a.shape[3] != self.cfg.num_keypoints:
            raise ValueError(f"poses must be (B,{self.cfg.in_channels},T,{self.cfg.num_keypoints}), got {a.shape}")
        B, _, T, _ = a.shape
        key = (T, chunk)
        if key not in self._runners:
            h = C.c_void_p()
            N.check(self._lib.sf_runner_create(self._h, T, chunk, C.byref(h)), "sf_runner_create")
            self._runners[key] = h
        out = np.empty(B, dtype=np.float32)
        r = self._runners[key]
        self._run_prec("score", T, precision, lambda p: N.check(
            self._lib.sf_runner_score(r, C.c_void_p(a.ctypes.data), B, p, C.c_void_p(out.ctypes.data)), "sf_runner_score"))
        return out

    def _runner(self, T: int, chunk: int) -> C.c_void_p:
        key = (T, chunk)
        if key not in self._runners:
            h = C.c_void_p()
            N.check(self._lib.sf_runner_create(self._h, T, chunk, C.byref(h)), "sf_runner_create")
            self._runners[key] = h
        return self._runners[key]

    # -- packed tracks in, scores + window index out (dataset construction + scoring loop of the reference in one call)
    def score_tracks(self, tracks: "DeviceTracks", seq_len: int, stride: int, max_gap: int = 5, normalize: bool = True,
                     add_neck: Optional[bool] = None, precision: str = "auto") -> Dict[str, torch.Tensor]:
        """Device-resident tracks -> per-window scores, labels and (track, start) index in the reference's window order
        (``sf_score_from_tracks``): the normalised windows only ever exist one 131,072-window pass at a time."""
        if tracks.device != self.device:
            raise ValueError(f"tracks live on {tracks.device}, the engine on {self.device}")
        nt = tracks.native()
        p = _window_params(seq_len, stride, max_gap, self.cfg.num_keypoints, normalize, add_neck, self.cfg.in_channels == 3)
        cap = self._lib.sf_window_capacity(C.byref(nt), C.byref(p))
        N.check(int(min(cap, 0)), "sf_window_capacity")
        capn = max(int(cap), 1)
        dev = self.device
        scores = torch.empty(capn, dtype=torch.float32, device=dev)
        labels = torch.empty(capn, dtype=torch.int32, device=dev)
        wtrack = torch.empty(capn, dtype=torch.int32, device=dev)
        wstart = torch.empty(capn, dtype=torch.int32, device=dev)
        wsb = self._lib.sf_score_from_tracks_workspace_bytes(self._h, C.byref(nt), C.byref(p))
        N.check(int(min(wsb, 0)), "sf_score_from_tracks_workspace_bytes")
        key = ("tracks", torch.cuda.current_stream(dev).cuda_stream)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < wsb:
            ws = self._ws[key] = torch.empty(max(int(wsb), 1), dtype=torch.uint8, device=dev)
        n_host = C.c_int64(0)
        st = _stream_ptr(dev)
        self._run_prec("score", seq_len, precision, lambda pr: N.check(
            self._lib.sf_score_from_tracks(self._h, C.byref(nt), C.byref(p), pr, _ptr(scores), _ptr(labels), _ptr(wtrack), _ptr(wstart),
                                           C.byref(n_host), _ptr(ws), int(wsb), st), "sf_score_from_tracks"))
        n = int(n_host.value)
        return {"scores": scores[:n], "labels": labels[:n], "window_track": wtrack[:n], "window_start": wstart[:n], "n_windows": n,
                "capacity": int(cap)}

    def score_tracks_host(self, tracks: "PackedTracks", seq_len: int, stride: int, max_gap: int = 5, normalize: bool = True,
                          add_neck: Optional[bool] = None, precision: str = "auto", chunk: int = 16384) -> Dict[str, np.ndarray]:
        """Host-resident tracks -> host scores + window index (``sf_runner_score_tracks``): groups of whole tracks are uploaded
        on a copy stream while the previous group is windowed and scored."""
        kp = np.ascontiguousarray(tracks.kp, dtype=np.float32)
        fno = np.ascontiguousarray(tracks.frame_no, dtype=np.int32)
        off = np.ascontiguousarray(tracks.track_offsets, dtype=np.int64)
        vid = np.ascontiguousarray(tracks.track_video, dtype=np.int32)
        has_gt = tracks.gt is not None and tracks.gt_offsets is not None and len(tracks.gt) > 0
        gt = np.ascontiguousarray(tracks.gt, dtype=np.uint8) if has_gt else None
        gto = np.ascontiguousarray(tracks.gt_offsets, dtype=np.int64) if has_gt else None
        t = N.SfTracks()
        t.kp_dev, t.frame_no_dev = kp.ctypes.data, fno.ctypes.data           # HOST pointers for this entry point
        t.track_offsets_host, t.track_video_host = off.ctypes.data, vid.ctypes.data
        t.gt_dev = gt.ctypes.data if has_gt else 0
        t.gt_offsets_host = gto.ctypes.data if has_gt else 0
        t.n_frames, t.n_tracks = int(kp.shape[0]), int(len(off) - 1)
        t.n_videos = int(len(gto) - 1) if has_gt else 0
        t.kp_per_frame, t.kp_channels = int(kp.shape[1]), int(kp.shape[2])
        p = _window_params(seq_len, stride, max_gap, self.cfg.num_keypoints, normalize, add_neck, self.cfg.in_channels == 3)
        cap = self._lib.sf_window_capacity(C.byref(t), C.byref(p))
        N.check(int(min(cap, 0)), "sf_window_capacity")
        capn = max(int(cap), 1)
        scores = np.empty(capn, np.float32)
        labels = np.empty(capn, np.int32)
        wtrack = np.empty(capn, np.int32)
        wstart = np.empty(capn, np.int32)
        n_host = C.c_int64(0)
        r = self._runner(seq_len, chunk)
        self._run_prec("score", seq_len, precision, lambda pr: N.check(
            self._lib.sf_runner_score_tracks(r, C.byref(t), C.byref(p), pr, C.c_void_p(scores.ctypes.data), C.c_void_p(labels.ctypes.data),
                                             C.c_void_p(wtrack.ctypes.data), C.c_void_p(wstart.ctypes.data), C.byref(n_host)),
            "sf_runner_score_tracks"))
        n = int(n_host.value)
        return {"scores": scores[:n], "labels": labels[:n], "window_track": wtrack[:n], "window_start": wstart[:n], "n_windows": n,
                "capacity": int(cap)}


# --------------------------------------------------------------------------- pose decoder (optional output)
class DecoderEngine:
    """Packed GCAE pose decoder on one GPU (``sf_decoder``): tokens (B,S,latent*V) -> poses (B,C,seq_len,V) in eval mode.
    Built from the decoder module's own ``state_dict()`` (BatchNorm folded natively)."""

    def __init__(self, latent_channels: int, hidden_channels: int, out_channels: int, num_keypoints: int, seq_len: int,
                 upsample: Sequence[int], state_dict: Dict[str, torch.Tensor], device: torch.device):
        lib = N.load()
        N.check(lib.sf_device_count(), "sf_device_count")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DecoderEngine needs a CUDA device; there is no CPU path")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.out_shape = (out_channels, seq_len, num_keypoints)
        self.d_tok = latent_channels * num_keypoints
        names, arrays = [], []
        for k, v in state_dict.items():
            if torch.is_tensor(v) and v.is_floating_point():
                names.append(k.encode())
                arrays.append(np.ascontiguousarray(v.detach().to("cpu", torch.float32).numpy()))
        n = len(names)
        ups = (C.c_int32 * len(upsample))(*[int(u) for u in upsample])
        h = C.c_void_p()
        N.check(lib.sf_decoder_create(latent_channels, hidden_channels, out_channels, num_keypoints, seq_len, len(upsample), ups, n,
                                      (C.c_char_p * n)(*names), (C.c_void_p * n)(*[a.ctypes.data for a in arrays]),
                                      (C.c_int64 * n)(*[a.size for a in arrays]), idx, C.byref(h)), "sf_decoder_create")
        self._lib, self._h = lib, h
        self._ws: Optional[torch.Tensor] = None

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.sf_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, tokens: torch.Tensor) -> torch.Tensor:
        if tokens.dim() != 3 or tokens.shape[2] != self.d_tok:
            raise ValueError(f"tokens must be (B,S,{self.d_tok}), got {tuple(tokens.shape)}")
        if not tokens.is_cuda:
            raise RuntimeError("native decoder takes CUDA tensors only (no CPU fallback)")
        t = tokens.to(self.device, torch.float32).contiguous()
        B, S, _ = t.shape
        out = torch.empty((B,) + self.out_shape, dtype=torch.float32, device=self.device)
        nb = self._lib.sf_decoder_workspace_bytes(self._h, B, S)
        N.check(int(min(nb, 0)), "sf_decoder_workspace_bytes")
        if self._ws is None or self._ws.numel() < nb:
            self._ws = torch.empty(max(int(nb), 1), dtype=torch.uint8, device=self.device)
        N.check(self._lib.sf_decode_poses(self._h, _ptr(t), B, S, _ptr(out), _ptr(self._ws), int(nb), _stream_ptr(self.device)),
                "sf_decode_poses")
        return out


# --------------------------------------------------------------------------- windowing
def _window_params(seq_len: int, stride: int, max_gap: int, num_keypoints: int, normalize: bool, add_neck: Optional[bool],
                   include_confidence: bool) -> N.SfWindowParams:
    p = N.SfWindowParams()
    p.seq_len, p.stride, p.max_gap, p.num_keypoints, p.normalize = seq_len, stride, max_gap, num_keypoints, int(normalize)
    # default = the shopformer_2 convention (18 keypoints = COCO-17 + synthetic neck); shopformer/ passes add_neck=False
    p.add_neck = int(num_keypoints == 18 if add_neck is None else add_neck)
    p.include_confidence = int(include_confidence)
    return p


@dataclass
class PackedTracks:
    """Packed per-person tracks (see include/shopformer_b200.h, ``sf_tracks``)."""
    kp: np.ndarray                 # (F, K, 3) fp32 (x, y, conf), or (F, K, 2) when the confidence was dropped at ingest
    frame_no: np.ndarray           # (F,) int32
    track_offsets: np.ndarray      # (n_tracks+1,) int64
    track_video: np.ndarray        # (n_tracks,) int32
    gt: Optional[np.ndarray] = None        # concatenated uint8
    gt_offsets: Optional[np.ndarray] = None  # (n_videos+1,) int64
    video_names: List[str] = field(default_factory=list)

    @property
    def n_tracks(self) -> int:
        return len(self.track_offsets) - 1


class DeviceTracks:
    """PackedTracks resident in HBM (upload once, window many times)."""

    def __init__(self, tracks: PackedTracks, device: torch.device):
        self.host = tracks
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.kp = torch.from_numpy(np.ascontiguousarray(tracks.kp, dtype=np.float32)).to(self.device)
        self.frame_no = torch.from_numpy(np.ascontiguousarray(tracks.frame_no, dtype=np.int32)).to(self.device)
        self.gt = None
        if tracks.gt is not None and tracks.gt_offsets is not None and len(tracks.gt) > 0:
            self.gt = torch.from_numpy(np.ascontiguousarray(tracks.gt, dtype=np.uint8)).to(self.device)
        self._off = np.ascontiguousarray(tracks.track_offsets, dtype=np.int64)
        self._vid = np.ascontiguousarray(tracks.track_video, dtype=np.int32)
        self._gto = None if tracks.gt_offsets is None else np.ascontiguousarray(tracks.gt_offsets, dtype=np.int64)

    def native(self) -> N.SfTracks:
        t = N.SfTracks()
        t.kp_dev = self.kp.data_ptr()
        t.frame_no_dev = self.frame_no.data_ptr()
        t.track_offsets_host = self._off.ctypes.data
        t.track_video_host = self._vid.ctypes.data
        t.gt_dev = 0 if self.gt is None else self.gt.data_ptr()
        t.gt_offsets_host = 0 if self._gto is None else self._gto.ctypes.data
        t.n_frames = int(self.kp.shape[0])
        t.n_tracks = int(len(self._off) - 1)
        t.n_videos = 0 if self._gto is None else int(len(self._gto) - 1)
        t.kp_per_frame = int(self.kp.shape[1])
        t.kp_channels = int(self.kp.shape[2])
        return t


def window_normalize(tracks: DeviceTracks, seq_len: int, stride: int, num_keypoints: int = 17, max_gap: int = 5,
                     normalize: bool = True, want_frame_indices: bool = False, sync: bool = True,
                     add_neck: Optional[bool] = None, include_confidence: bool = False):
    """Run the windowing kernels.  Returns a dict of CUDA tensors trimmed to the number of
    valid windows when ``sync`` (one 8-byte D2H), else capacity-sized tensors + ``n_windows``
    as a device scalar."""
    lib = N.load()
    dev = tracks.device
    nt = tracks.native()
    p = _window_params(seq_len, stride, max_gap, num_keypoints, normalize, add_neck, include_confidence)
    n_planes = 3 if include_confidence else 2
    cap = lib.sf_window_capacity(C.byref(nt), C.byref(p))
    N.check(int(min(cap, 0)), "sf_window_capacity")
    wsb = lib.sf_window_workspace_bytes(C.byref(nt), C.byref(p))
    N.check(int(min(wsb, 0)), "sf_window_workspace_bytes")
    capn = max(int(cap), 1)
    poses = torch.empty(capn, n_planes, seq_len, num_keypoints, dtype=torch.float32, device=dev)
    labels = torch.empty(capn, dtype=torch.int32, device=dev)
    wtrack = torch.empty(capn, dtype=torch.int32, device=dev)
    wstart = torch.empty(capn, dtype=torch.int32, device=dev)
    fidx = torch.empty(capn, seq_len, dtype=torch.int32, device=dev) if want_frame_indices else None
    n_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(max(int(wsb), 1), dtype=torch.uint8, device=dev)
    n_host = C.c_int64(0)
    rc = lib.sf_window_normalize(C.byref(nt), C.byref(p), _ptr(poses), _ptr(labels), _ptr(wtrack), _ptr(wstart), _ptr(fidx),
                                 _ptr(n_dev), C.byref(n_host) if sync else None, _ptr(ws), int(wsb), _stream_ptr(dev))
    N.check(rc, "sf_window_normalize")
    out = {"poses": poses, "labels": labels, "window_track": wtrack, "window_start": wstart, "frame_indices": fidx,
           "n_windows": n_dev, "capacity": int(cap)}
    if sync:
        n = int(n_host.value)
        for k in ("poses", "labels", "window_track", "window_start", "frame_indices"):
            if out[k] is not None:
                out[k] = out[k][:n]
        out["n_windows"] = n
    return out


# --------------------------------------------------------------------------- after the path: aggregation + ranking metrics
def video_aggregate(scores: torch.Tensor, video_ids: torch.Tensor, labels: Optional[torch.Tensor], n_videos: int) -> Dict[str, torch.Tensor]:
    """Per-video max / mean / 95th percentile (float64, numpy semantics) of window scores, the video label (label of the
    video's last window, as shopformer_2/evaluate.py:107-116 leaves it) and the window count, on the device
    (``sf_video_aggregate``).  Videos without windows get NaN / label 0 / count 0."""
    lib = N.load()
    if not scores.is_cuda:
        raise RuntimeError("video_aggregate takes CUDA tensors (no CPU fallback)")
    dev = scores.device
    s = scores.detach().to(torch.float32).contiguous().view(-1)
    v = video_ids.to(dev, torch.int32).contiguous().view(-1)
    lb = None if labels is None else labels.to(dev, torch.int32).contiguous().view(-1)
    n = int(s.numel())
    if v.numel() != n or (lb is not None and lb.numel() != n):
        raise ValueError("scores, video_ids and labels must have the same length")
    nv = int(n_videos)
    out = {k: torch.empty(max(nv, 1), dtype=torch.float64, device=dev) for k in ("max", "mean", "percentile_95")}
    vlabel = torch.zeros(max(nv, 1), dtype=torch.int32, device=dev)
    count = torch.zeros(max(nv, 1), dtype=torch.int32, device=dev)
    wsb = lib.sf_video_aggregate_workspace_bytes(n, nv)
    N.check(int(min(wsb, 0)), "sf_video_aggregate_workspace_bytes")
    ws = torch.empty(max(int(wsb), 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(lib.sf_video_aggregate(_ptr(s), _ptr(v), _ptr(lb), n, nv, _ptr(out["max"]), _ptr(out["mean"]), _ptr(out["percentile_95"]),
                                       _ptr(vlabel), _ptr(count), _ptr(ws), int(wsb), _stream_ptr(dev)), "sf_video_aggregate")
    res = {k: t[:nv] for k, t in out.items()}
    res["label"] = vlabel[:nv]
    res["count"] = count[:nv]
    return res


def ranking_metrics(scores: torch.Tensor, labels: torch.Tensor, threshold: Optional[float] = None) -> Dict[str, float]:
    """AUC-ROC, average precision, the threshold (given, or the Youden-J optimum of the ROC curve like the reference's
    ``compute_metrics``) and the confusion counts at it, computed on the device (``sf_ranking_metrics``: radix sort + scans).
    Degenerate label sets follow the reference's conventions (AUC-ROC 0.5, AUC-PR 0.0)."""
    lib = N.load()
    if not scores.is_cuda:
        raise RuntimeError("ranking_metrics takes CUDA tensors (no CPU fallback)")
    dev = scores.device
    s = scores.detach().to(torch.float32).contiguous().view(-1)
    lb = labels.to(dev, torch.int32).contiguous().view(-1)
    n = int(s.numel())
    if lb.numel() != n or n == 0:
        raise ValueError("scores and labels must be non-empty and of the same length")
    wsb = lib.sf_ranking_metrics_workspace_bytes(n)
    N.check(int(min(wsb, 0)), "sf_ranking_metrics_workspace_bytes")
    ws = torch.empty(int(wsb), dtype=torch.uint8, device=dev)
    out = (C.c_double * 8)()
    thr = float("nan") if threshold is None else float(threshold)
    with torch.cuda.device(dev):
        N.check(lib.sf_ranking_metrics(_ptr(s), _ptr(lb), n, C.c_float(thr), out, _ptr(ws), int(wsb), _stream_ptr(dev)), "sf_ranking_metrics")
    auc_roc, auc_pr, thr_used, n_pos, tp, fp, tn, fn = [float(x) for x in out]
    if auc_roc != auc_roc:
        auc_roc = 0.5
    if auc_pr != auc_pr:
        auc_pr = 0.0
    precision = tp / (tp + fp) if tp + fp > 0 else 0.0
    recall = tp / (tp + fn) if tp + fn > 0 else 0.0
    f1 = 2 * precision * recall / (precision + recall) if precision + recall > 0 else 0.0
    return {"auc_roc": auc_roc, "auc_pr": auc_pr, "accuracy": (tp + tn) / n, "precision": precision, "recall": recall, "f1": f1,
            "threshold": thr_used, "tp": int(tp), "fp": int(fp), "tn": int(tn), "fn": int(fn), "n_positive": int(n_pos)}
