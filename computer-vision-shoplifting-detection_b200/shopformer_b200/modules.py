"""Parameter containers + autograd (training) path shared by the two drop-in packages.

The module tree reproduces the reference's attribute names exactly so that
``state_dict()`` keys, shapes and dtypes are interchangeable with reference checkpoints
 and so that seeded construction consumes torch's RNG in the same order.

Only TRAINING (``module.training`` or grad enabled) runs through these ``forward``
bodies -- they are ordinary ATen compositions so autograd, train-mode BatchNorm and
dropout behave as in the reference.  Eval-mode inference on CUDA tensors is routed by the
facades (``shopformer/models/shopformer.py`` and ``shopformer_2/models/shopformer.py`` in
this tree) to the native sm_100a kernels via :mod:`shopformer_b200.engine`; eval-mode
inference on CPU tensors raises -- there is no CPU inference path.

Reference for the maths: shopformer/models/gcae.py:19-549, shopformer/models/transformer.py:14-349,
shopformer_2/models/gcae.py:22-613, shopformer_2/models/transformer.py:18-224.
"""
from __future__ import annotations

import math
import os
import weakref
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import model_handle

# --------------------------------------------------------------------------- skeleton graphs
_COCO17 = [(0, 1), (0, 2), (1, 3), (2, 4), (0, 5), (0, 6), (5, 7), (7, 9), (6, 8), (8, 10),
           (5, 11), (6, 12), (11, 12), (11, 13), (13, 15), (12, 14), (14, 16)]
_OPENPOSE18 = [(0, 1), (0, 14), (0, 15), (14, 16), (15, 17), (1, 2), (2, 3), (3, 4), (1, 5), (5, 6),
               (6, 7), (1, 8), (8, 9), (9, 10), (1, 11), (11, 12), (12, 13)]
_COCO_NECK18 = [(0, 1), (0, 2), (1, 3), (2, 4), (0, 17), (17, 5), (17, 6), (5, 7), (7, 9), (6, 8),
                (8, 10), (5, 11), (6, 12), (11, 12), (11, 13), (13, 15), (12, 14), (14, 16)]


def skeleton_edges(num_keypoints: int, layout: str, family: int) -> List[Tuple[int, int]]:
    """Edge list selection rule of each variant.
    family 1: shopformer/models/gcae.py:30-56 (layout decides; unknown layout -> ValueError).
    family 2: shopformer_2/models/gcae.py:33-68 ('coco' & 17 -> COCO; 18 keypoints or
    'coco_with_neck' -> COCO+neck; otherwise ValueError)."""
    if family == 1:
        if layout == "coco":
            return _COCO17
        if layout == "openpose":
            return _OPENPOSE18
        raise ValueError(f"Unknown skeleton layout: {layout}")
    if layout == "coco" and num_keypoints == 17:
        return _COCO17
    if num_keypoints == 18 or layout == "coco_with_neck":
        return _COCO_NECK18
    raise ValueError(f"Unknown layout: {layout} with {num_keypoints} keypoints")


def get_skeleton_adjacency(num_keypoints: int = 17, layout: str = "coco", family: int = 1) -> np.ndarray:
    """A + I as float64 (edges touching a keypoint index >= V are dropped)."""
    a = np.zeros((num_keypoints, num_keypoints))
    for i, j in skeleton_edges(num_keypoints, layout, family):
        if i < num_keypoints and j < num_keypoints:
            a[i, j] = a[j, i] = 1
    return a + np.eye(num_keypoints)


def normalize_adjacency(adj: np.ndarray) -> np.ndarray:
    """D^-1/2 (A) D^-1/2 in float64 (gcae.py:71-85)."""
    deg = adj.sum(axis=1)
    with np.errstate(divide="ignore"):
        dis = np.power(deg, -0.5)
    dis[np.isinf(dis)] = 0.0
    return np.diag(dis) @ adj @ np.diag(dis)


def strides_halving(seq_len: int, num_tokens: int, num_layers: int) -> List[int]:
    """variant 1 stride rule (shopformer/models/gcae.py:317-329)."""
    out, cur, i = [1] * num_layers, seq_len, 0
    while cur > num_tokens and i < num_layers:
        if cur // 2 >= num_tokens:
            out[i], cur = 2, cur // 2
        i += 1
    return out


def strides_factorised(seq_len: int, num_tokens: int, num_layers: int) -> Tuple[List[int], bool]:
    """variant 2 stride rule + pooling flag (shopformer_2/models/gcae.py:331-373)."""
    out = [1] * num_layers
    rem, fac = seq_len // num_tokens, []
    for p in (2, 3, 4, 5, 6):
        while rem % p == 0 and rem > 1:
            fac.append(p)
            rem //= p
    if rem > 1:
        fac.append(rem)
    for i, f in enumerate(sorted(fac)):
        if i < num_layers:
            out[i] = f
    out.sort(reverse=True)
    n = seq_len
    for s in out:
        n //= s
    return out, n != num_tokens


def upsample_factors(num_tokens: int, seq_len: int, num_layers: int) -> List[int]:
    """decoder doubling rule (gcae.py:437-449)."""
    out, cur, i = [1] * num_layers, num_tokens, 0
    while cur < seq_len and i < num_layers:
        if cur * 2 <= seq_len:
            out[i], cur = 2, cur * 2
        i += 1
    return out


# --------------------------------------------------------------------------- dispatch helpers
def composite_eval_allowed() -> bool:
    """Test-only escape hatch for GPU-less hosts (exercising host logic of the facades)."""
    return os.environ.get("SHOPFORMER_B200_COMPOSITE_EVAL", "0") == "1"


def eager_baseline_forced() -> bool:
    """Measurement-only switch (bench.py's `eager_aten_gpu` leg): eval-mode inference runs the ATen composition of the
    same modules on the GPU -- what the reference's eager PyTorch does on the same device -- instead of the native kernels."""
    return os.environ.get("SHOPFORMER_B200_EAGER_BASELINE", "0") == "1"


def wants_native(module: nn.Module, x: torch.Tensor) -> bool:
    """True when this call is eval-mode inference and must run on the sm_100a kernels."""
    if module.training or torch.is_grad_enabled():
        return False                      # training / autograd: ATen composition below
    if x.is_cuda and eager_baseline_forced():
        return False
    owner_ref = getattr(module, "_sf_owner", None)
    owner = owner_ref() if owner_ref is not None else None
    if owner is not None and owner.training:
        # stage-2 training with a frozen, eval-mode tokenizer (shopformer_2 freeze_gcae): the facade's other parameters
        # change every optimizer step, so the facade's packed model would be rebuilt every step.  The tokenizer itself
        # still runs on the native kernels when it is the caller (`_sf_tok_only`): it gets its own packed model keyed on
        # the encoder's tensors alone (first item).
        if not (getattr(module, "_sf_tok_only", False) and x.is_cuda and all(not p.requires_grad for p in module.parameters())):
            return False
        return True
    if x.is_cuda:
        return True
    if composite_eval_allowed():
        return False
    raise RuntimeError(
        "shopformer_b200: eval-mode inference runs on CUDA (sm_100a) only and there is no CPU fallback; "
        "move the model and the input to a B200 (`.to('cuda')`).")


class _Owned:
    """Mixin: sub-modules reach the facade's packed-weight engine through a weak reference."""
    _sf_owner: Optional[Callable[[], Optional[nn.Module]]] = None

    def _engine(self):
        owner = self._sf_owner() if self._sf_owner is not None else None
        if owner is None:
            return None
        if owner.training and getattr(self, "_sf_tok_only", False):
            return owner._sf_tok_engine()          # frozen tokenizer inside a training facade
        return owner._sf_engine()

    def _precision(self) -> str:
        owner = self._sf_owner() if self._sf_owner is not None else None
        return owner._sf_resolve_precision() if owner is not None else "auto"


def adopt(owner: nn.Module, *children: nn.Module) -> None:
    ref = weakref.ref(owner)
    for ch in children:
        object.__setattr__(ch, "_sf_owner", ref)


# --------------------------------------------------------------------------- ST-GCN pieces
class GraphConvolution(nn.Module):
    """Y = A_hat X W + b over the keypoint graph (parameters: weight (Cin,Cout), bias; buffer adj)."""

    def __init__(self, in_channels: int, out_channels: int, adj: torch.Tensor, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.register_buffer("adj", adj)
        self.weight = nn.Parameter(torch.FloatTensor(in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.FloatTensor(out_channels))
        else:
            self.register_parameter("bias", None)
        nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:          # (B,C,T,V) -> (B,Cout,T,V)
        y = torch.einsum("vu,bctu->bctv", self.adj, x)
        y = torch.einsum("bctv,co->botv", y, self.weight)
        return y if self.bias is None else y + self.bias.view(1, -1, 1, 1)


class TemporalConvolution(nn.Module):
    """(k x 1) conv along time + BatchNorm2d."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 9, stride: int = 1,
                 padding: int = 4, dilation: int = 1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=(kernel_size, 1), stride=(stride, 1),
                              padding=(padding, 0), dilation=(dilation, 1))
        self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.bn(self.conv(x))


class STGCNBlock(nn.Module):
    """relu(dropout(tcn(relu(gcn(x)))) + residual(x))."""

    def __init__(self, in_channels: int, out_channels: int, adj: torch.Tensor, stride: int = 1,
                 residual: bool = True, dropout: float = 0.0):
        super().__init__()
        self.gcn = GraphConvolution(in_channels, out_channels, adj)
        self.tcn = TemporalConvolution(out_channels, out_channels, kernel_size=9, stride=stride, padding=4)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(dropout)
        self.stride = stride
        if not residual:
            self.residual = lambda x: 0
        elif in_channels == out_channels and stride == 1:
            self.residual = lambda x: x
        else:
            self.residual = nn.Sequential(
                nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=(stride, 1)),
                nn.BatchNorm2d(out_channels))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        skip = self.residual(x)
        y = self.dropout(self.tcn(self.relu(self.gcn(x))))
        return self.relu(y + skip)


class GCAEEncoder(nn.Module, _Owned):
    """Pose window (B,C,T,V) -> tokens (B,S,Cout*V).  ``family`` picks the stride/graph rules."""
    _sf_tok_only = True        # a frozen, eval-mode tokenizer inside a training facade gets its own packed model

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_keypoints: int,
                 seq_len: int, num_tokens: int, num_layers: int, dropout: float, layout: str, family: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_keypoints, self.seq_len, self.num_tokens = num_keypoints, seq_len, num_tokens
        self._family = family
        adj = torch.FloatTensor(normalize_adjacency(get_skeleton_adjacency(num_keypoints, layout, family)))
        self.bn_input = nn.BatchNorm1d(in_channels * num_keypoints)
        chans = [in_channels, hidden_channels, hidden_channels, hidden_channels, out_channels]
        if family == 1:
            self.strides = strides_halving(seq_len, num_tokens, num_layers)
            self._needs_pooling = False
        else:
            self.strides, self._needs_pooling = strides_factorised(seq_len, num_tokens, num_layers)
        self.layers = nn.ModuleList([
            STGCNBlock(chans[i], chans[i + 1], adj, stride=self.strides[i], residual=True, dropout=dropout)
            for i in range(num_layers)])
        if family == 2:
            self.adaptive_pool = nn.AdaptiveAvgPool2d((num_tokens, num_keypoints))
        self._channels = chans[:num_layers + 1]

    def _as_bctv(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 4 and x.shape[-1] == self.in_channels:      # (B,T,V,C) input
            x = x.permute(0, 3, 1, 2).contiguous()
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self._as_bctv(x)
        if wants_native(self, x):
            eng = self._engine()
            if eng is not None:
                return torch.ops.shopformer_b200.tokenize(x, model_handle(eng), self._precision())
        b, c, t, v = x.shape
        y = self.bn_input(x.permute(0, 1, 3, 2).reshape(b, c * v, t))
        x = y.view(b, c, v, t).permute(0, 1, 3, 2).contiguous()
        for blk in self.layers:
            x = blk(x)
        if self._needs_pooling:
            x = self.adaptive_pool(x)
        b, c, t, v = x.shape
        return x.permute(0, 2, 1, 3).reshape(b, t, c * v)


class GCAEDecoder(nn.Module):
    """tokens (B,S,D) -> poses (B,C,T,V): Linear, then (ConvTranspose|Conv1x1)+BN+ReLU stack."""

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_keypoints: int,
                 seq_len: int, num_tokens: int, num_layers: int, dropout: float, layout: str, family: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_keypoints, self.seq_len, self.num_tokens = num_keypoints, seq_len, num_tokens
        if family == 1:                                           # validates `layout` like the reference does
            skeleton_edges(num_keypoints, layout, family)
        self.initial_proj = nn.Linear(in_channels * num_keypoints, hidden_channels * num_keypoints)
        ups = upsample_factors(num_tokens, seq_len, num_layers)
        self._sf_hidden, self._sf_upsample = hidden_channels, list(ups)
        outs = [hidden_channels] * (num_layers - 1) + [out_channels]
        seq: List[nn.Module] = []
        for i, (u, oc) in enumerate(zip(ups, outs)):
            if u > 1:
                seq.append(nn.ConvTranspose2d(hidden_channels, oc, kernel_size=(u, 1), stride=(u, 1)))
            else:
                seq.append(nn.Conv2d(hidden_channels, oc, kernel_size=1))
            if i < num_layers - 1:
                seq += [nn.BatchNorm2d(oc), nn.ReLU(inplace=True), nn.Dropout(dropout)]
        self.layers = nn.Sequential(*seq)

    def _sf_decoder_engine(self):
        """Packed native decoder (sf_decoder), rebuilt when a parameter / buffer changed or moved."""
        from .engine import DecoderEngine
        tensors = list(self.parameters()) + list(self.buffers())
        fp = tuple((t.data_ptr(), t._version) for t in tensors)
        cached = self.__dict__.get("_sf_dec_cache")
        if cached is not None and cached[0] == fp:
            return cached[1]
        if cached is not None:
            cached[1].close()
        eng = DecoderEngine(self.in_channels, self._sf_hidden, self.out_channels, self.num_keypoints, self.seq_len, self._sf_upsample,
                            self.state_dict(), self.initial_proj.weight.device)
        self.__dict__["_sf_dec_cache"] = (fp, eng)
        return eng

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        if z.dim() == 3 and wants_native(self, z):
            from .native import NativeError
            try:
                return self._sf_decoder_engine().decode(z)
            except NativeError as exc:
                if exc.code != -4:                    # SF_E_UNSUPPORTED (a window's activations exceed shared memory): ATen below
                    raise
        b, s, _ = z.shape
        h = self.initial_proj(z)
        h = h.view(b, s, h.shape[-1] // self.num_keypoints, self.num_keypoints).permute(0, 2, 1, 3).contiguous()
        h = self.layers(h)
        if h.shape[2] != self.seq_len:
            h = F.interpolate(h, size=(self.seq_len, self.num_keypoints), mode="bilinear", align_corners=False)
        return h


class GCAE(nn.Module):
    """Graph-convolutional auto-encoder = tokenizer (encoder) + pose decoder."""

    def __init__(self, in_channels: int, hidden_channels: int, latent_channels: int, num_keypoints: int,
                 seq_len: int, num_tokens: int, num_layers: int, dropout: float, layout: str, family: int):
        super().__init__()
        self.in_channels, self.num_keypoints = in_channels, num_keypoints
        self.seq_len, self.num_tokens = seq_len, num_tokens
        kw = dict(num_keypoints=num_keypoints, seq_len=seq_len, num_tokens=num_tokens, num_layers=num_layers,
                  dropout=dropout, layout=layout, family=family)
        self.encoder = GCAEEncoder(in_channels, hidden_channels, latent_channels, **kw)
        self.decoder = GCAEDecoder(latent_channels, hidden_channels, in_channels, **kw)
        self.embedding_dim = latent_channels * num_keypoints

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        return self.encoder(x)

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        return self.decoder(z)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        tokens = self.encode(x)
        return self.decode(tokens), tokens

    def get_embedding_dim(self) -> int:
        return self.embedding_dim


# --------------------------------------------------------------------------- transformer pieces
def sinusoid_table(d_model: int, max_len: int) -> torch.Tensor:
    """(1, max_len, d_model) fp32 table: sin on even, cos on odd feature indices."""
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * (div if d_model % 2 == 0 else div[:-1]))
    return pe.unsqueeze(0)


class PostLNEncoderLayer(nn.Module):
    """x = LN(x + SA(x)); x = LN(x + FFN(x))  (variant 1)."""

    def __init__(self, d_model: int, nhead: int, dim_feedforward: int = 64, dropout: float = 0.1):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.activation = nn.ReLU()

    def forward(self, src, src_mask=None, src_key_padding_mask=None):
        a, _ = self.self_attn(src, src, src, attn_mask=src_mask, key_padding_mask=src_key_padding_mask)
        src = self.norm1(src + self.dropout1(a))
        f = self.linear2(self.dropout(self.activation(self.linear1(src))))
        return self.norm2(src + self.dropout2(f))


class PostLNDecoderLayer(nn.Module):
    """x = LN(x + SA(x)); x = LN(x + CA(x, mem)); x = LN(x + FFN(x))  (variant 1)."""

    def __init__(self, d_model: int, nhead: int, dim_feedforward: int = 64, dropout: float = 0.1):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.activation = nn.ReLU()

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None):
        a, _ = self.self_attn(tgt, tgt, tgt, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask)
        tgt = self.norm1(tgt + self.dropout1(a))
        a, _ = self.multihead_attn(tgt, memory, memory, attn_mask=memory_mask,
                                   key_padding_mask=memory_key_padding_mask)
        tgt = self.norm2(tgt + self.dropout2(a))
        f = self.linear2(self.dropout(self.activation(self.linear1(tgt))))
        return self.norm3(tgt + self.dropout3(f))
