"""Streaming mode (BASELINE config #5): N concurrent camera/person streams, one new window per stream per tick.

Every stream owns a ring of its last T detections in HBM.  A tick appends `stride` new detections per stream
(one H2D copy from pinned memory), then runs normalise -> tokenizer -> transformer -> score over one window per
stream with fixed shapes, so the whole device side of a tick is captured ONCE in a CUDA graph and replayed
(launch-bound inner loop).  Reference analogue: `predict_poses` called per sample
(shopformer/inference.py:67-94, 145-151).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import native as N
from .engine import ScoringEngine, _ptr, _stream_ptr


class StreamScorer:
    def __init__(self, engine: ScoringEngine, n_streams: int, seq_len: int, stride: int, kp_per_frame: int = 17,
                 normalize: bool = True, precision: str = "fp32", use_graph: bool = True):
        if not 1 <= stride <= seq_len:
            raise ValueError("stride must be in [1, seq_len]")
        self.eng, self.n, self.T, self.stride, self.K = engine, n_streams, seq_len, stride, kp_per_frame
        self.V = engine.cfg.num_keypoints
        self.normalize, self.precision = normalize, precision
        dev = engine.device
        self._lib = N.load()
        # two rings (ping-pong) so that the shift is an out-of-place copy
        self.ring = [torch.zeros(n_streams, seq_len, kp_per_frame, 3, device=dev) for _ in range(2)]
        self.cur = 0
        self.new_dev = torch.zeros(n_streams, stride, kp_per_frame, 3, device=dev)
        self.new_pin = torch.zeros(n_streams, stride, kp_per_frame, 3).pin_memory()
        self.poses = torch.empty(n_streams, 2, seq_len, self.V, device=dev)
        self.scores = torch.empty(n_streams, device=dev)
        self.scores_pin = torch.empty(n_streams).pin_memory()
        self.stream = torch.cuda.Stream(device=dev)
        self.graphs = [None, None] if use_graph else None
        self.ticks = 0

    # device side of one tick, reading ring[src] and producing ring[dst]
    def _device_tick(self, src: int, dst: int) -> None:
        s, T = self.stride, self.T
        if s < T:
            self.ring[dst][:, :T - s].copy_(self.ring[src][:, s:])
        self.ring[dst][:, T - s:].copy_(self.new_dev)
        rc = self._lib.sf_normalize_windows(_ptr(self.ring[dst]), self.n, T, self.K, self.V, int(self.normalize),
                                            _ptr(self.poses), _stream_ptr(self.eng.device))
        N.check(rc, "sf_normalize_windows")
        self.eng.score_windows(self.poses, precision=self.precision, out=self.scores)

    def tick(self, new_frames: np.ndarray) -> np.ndarray:
        """new_frames: (n_streams, stride, K, 3) fp32 -> scores (n_streams,) of the windows ending at these frames."""
        self.new_pin.copy_(torch.from_numpy(np.ascontiguousarray(new_frames, dtype=np.float32)))
        src, dst = self.cur, self.cur ^ 1
        with torch.cuda.stream(self.stream):
            self.new_dev.copy_(self.new_pin, non_blocking=True)
            if self.graphs is None:
                self._device_tick(src, dst)
            else:
                if self.graphs[src] is None:
                    self._device_tick(src, dst)                      # warm-up run (allocations, attributes) outside capture
                    self.stream.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.stream):
                        self._device_tick(src, dst)
                    self.graphs[src] = g
                else:
                    self.graphs[src].replay()
            self.scores_pin.copy_(self.scores, non_blocking=True)
        self.stream.synchronize()
        self.cur = dst
        self.ticks += 1
        return self.scores_pin.numpy().copy()
