"""CPU oracle for the Shopformer scoring path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A functional (state-dict in, tensors out) restatement of the reference's eval-mode
maths, written with plain torch CPU ops so that it runs in fp32 *and* fp64.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file; the product path never does.

Parity status: **pinned**.  ``tests/golden/make_golden.py`` imports the real reference
from ``/root/reference`` in the build container, runs it on seeded inputs/weights and
commits its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
every function below against those vectors (fp64 to ~1e-12, fp32 to ~1e-5).

Each function cites the reference lines it restates (paths relative to the
reference checkout).  Two variants exist side by side in the reference:

  v1 = ``shopformer/``   post-LN, ReLU, shifted decoder target, score vs tokens+PE
  v2 = ``shopformer_2/`` pre-LN, GELU, final LayerNorms, optional 136<->144
       projections, score vs raw tokens
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5
LN_EPS = 1e-5


# --------------------------------------------------------------------------- config helpers
def strides_v1(seq_len: int, num_tokens: int, num_layers: int = 4) -> List[int]:
    """shopformer/models/gcae.py:317-329 -- greedy halving while len//2 >= num_tokens."""
    strides = [1] * num_layers
    cur, i = seq_len, 0
    while cur > num_tokens and i < num_layers:
        if cur // 2 >= num_tokens:
            strides[i] = 2
            cur //= 2
        i += 1
    return strides


def strides_v2(seq_len: int, num_tokens: int, num_layers: int = 4) -> List[int]:
    """shopformer_2/models/gcae.py:331-373 -- factorise seq_len//num_tokens over
    [2,3,4,5,6], ascending onto the first layers, then the whole list sorted descending."""
    strides = [1] * num_layers
    rem = seq_len // num_tokens
    factors: List[int] = []
    for p in (2, 3, 4, 5, 6):
        while rem % p == 0 and rem > 1:
            factors.append(p)
            rem //= p
    if rem > 1:
        factors.append(rem)
    factors.sort()
    for i, f in enumerate(factors):
        if i < num_layers:
            strides[i] = f
    strides.sort(reverse=True)
    return strides


def v2_needs_pool(seq_len: int, num_tokens: int, strides: Sequence[int]) -> bool:
    """shopformer_2/models/gcae.py:365-371 -- pooling flag uses FLOOR division lengths."""
    n = seq_len
    for s in strides:
        n //= s
    return n != num_tokens


def conv_len(t: int, s: int) -> int:
    """Conv2d k=9, pad=4, stride s along time: (t + 8 - 9)//s + 1."""
    return (t - 1) // s + 1


# --------------------------------------------------------------------------- building blocks
def _bn(x: Tensor, sd: SD, p: str, ch_dim: int) -> Tensor:
    """Eval-mode BatchNorm: (x - running_mean) * rsqrt(running_var + eps) * w + b."""
    shape = [1] * x.dim()
    shape[ch_dim] = -1
    rm, rv = sd[p + "running_mean"].to(x.dtype), sd[p + "running_var"].to(x.dtype)
    w, b = sd[p + "weight"].to(x.dtype), sd[p + "bias"].to(x.dtype)
    return (x - rm.view(shape)) * torch.rsqrt(rv.view(shape) + BN_EPS) * w.view(shape) + b.view(shape)


def _lin(x: Tensor, sd: SD, p: str) -> Tensor:
    return x @ sd[p + "weight"].to(x.dtype).t() + sd[p + "bias"].to(x.dtype)


def _ln(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"].to(x.dtype), sd[p + "bias"].to(x.dtype), LN_EPS)


def _mha(q_in: Tensor, kv_in: Tensor, sd: SD, p: str, nhead: int) -> Tensor:
    """nn.MultiheadAttention(batch_first=True), no masks, eval: packed in_proj split
    [0:d | d:2d | 2d:3d], contiguous head split, softmax(q k^T / sqrt(hd)) v, out_proj.
    Call sites: shopformer/models/transformer.py:108,180,186;
    shopformer_2/models/transformer.py:105-136 (nn.TransformerEncoder/DecoderLayer)."""
    d = q_in.shape[-1]
    w = sd[p + "in_proj_weight"].to(q_in.dtype)
    b = sd[p + "in_proj_bias"].to(q_in.dtype)
    q = q_in @ w[:d].t() + b[:d]
    k = kv_in @ w[d:2 * d].t() + b[d:2 * d]
    v = kv_in @ w[2 * d:].t() + b[2 * d:]
    B, Sq, _ = q.shape
    Sk = k.shape[1]
    hd = d // nhead
    q = q.view(B, Sq, nhead, hd).transpose(1, 2)
    k = k.view(B, Sk, nhead, hd).transpose(1, 2)
    v = v.view(B, Sk, nhead, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, Sq, d)
    return _lin(o, sd, p + "out_proj.")


# --------------------------------------------------------------------------- tokenizer (both variants)
def tokenize(sd: SD, poses: Tensor, strides: Sequence[int], *, prefix: str = "gcae.encoder.",
             pool_tokens: Optional[int] = None) -> Tensor:
    """GCAEEncoder.forward in eval mode.

    shopformer/models/gcae.py:331-366 (+ blocks 242-259, gcn 124-154, tcn 185-195);
    shopformer_2/models/gcae.py:375-422 (adaptive pool 406-415 when ``pool_tokens``).
    ``poses`` is (B,C,T,V) or (B,T,V,C) -- the latter detected exactly like the
    reference does: 4-D and last dim == in_channels.
    Returns tokens (B, T', Cout*V) with feature index c*V+v.
    """
    n_layers = len(strides)
    c_in = sd[prefix + "layers.0.gcn.weight"].shape[0]
    x = poses
    if x.dim() == 4 and x.shape[-1] == c_in:
        x = x.permute(0, 3, 1, 2)
    B, C, T, V = x.shape
    # E0: BatchNorm1d over channel index c*V+v, applied on (B, C*V, T)
    x = _bn(x.permute(0, 1, 3, 2).reshape(B, C * V, T), sd, prefix + "bn_input.", 1)
    x = x.view(B, C, V, T).permute(0, 1, 3, 2)
    for i in range(n_layers):
        p = f"{prefix}layers.{i}."
        s = int(strides[i])
        adj = sd[p + "gcn.adj"].to(x.dtype)
        w = sd[p + "gcn.weight"].to(x.dtype)
        # residual branch
        if (p + "residual.0.weight") in sd:
            r = F.conv2d(x, sd[p + "residual.0.weight"].to(x.dtype), sd[p + "residual.0.bias"].to(x.dtype),
                         stride=(s, 1))
            r = _bn(r, sd, p + "residual.1.", 1)
        else:
            r = x
        # spatial graph conv: Y[b,o,t,v] = sum_c (sum_u A[v,u] X[b,c,t,u]) W[c,o] + bias[o]
        g = torch.einsum("vu,bctu->bctv", adj, x)
        g = torch.einsum("bctv,co->botv", g, w) + sd[p + "gcn.bias"].to(x.dtype).view(1, -1, 1, 1)
        g = torch.relu(g)
        # temporal conv 9x1, stride (s,1), pad (4,0) + BN
        h = F.conv2d(g, sd[p + "tcn.conv.weight"].to(x.dtype), sd[p + "tcn.conv.bias"].to(x.dtype),
                     stride=(s, 1), padding=(4, 0))
        h = _bn(h, sd, p + "tcn.bn.", 1)
        x = torch.relu(h + r)
    if pool_tokens is not None:
        x = F.adaptive_avg_pool2d(x, (pool_tokens, V))
    B, C, T, V = x.shape
    return x.permute(0, 2, 1, 3).reshape(B, T, C * V)


# --------------------------------------------------------------------------- transformer v1
def reconstruct_v1(sd: SD, tokens: Tensor, nhead: int, *, prefix: str = "transformer.") -> Tensor:
    """ShopformerTransformer.forward, shopformer/models/transformer.py:304-329
    (encode 261-278, decode 280-302, layers 90-118 and 156-196). Post-LN, ReLU."""
    B, S, d = tokens.shape
    pe = sd[prefix + "pos_encoder.pe"].to(tokens.dtype)[:, :S]
    n_enc = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith(prefix + "encoder_layers."))
    n_dec = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith(prefix + "decoder_layers."))
    src = tokens + pe
    for i in range(n_enc):
        p = f"{prefix}encoder_layers.{i}."
        src = _ln(src + _mha(src, src, sd, p + "self_attn.", nhead), sd, p + "norm1.")
        ff = _lin(torch.relu(_lin(src, sd, p + "linear1.")), sd, p + "linear2.")
        src = _ln(src + ff, sd, p + "norm2.")
    mem = src
    tgt = torch.cat([torch.zeros(B, 1, d, dtype=tokens.dtype), tokens[:, :-1]], dim=1) + pe
    for i in range(n_dec):
        p = f"{prefix}decoder_layers.{i}."
        tgt = _ln(tgt + _mha(tgt, tgt, sd, p + "self_attn.", nhead), sd, p + "norm1.")
        tgt = _ln(tgt + _mha(tgt, mem, sd, p + "multihead_attn.", nhead), sd, p + "norm2.")
        ff = _lin(torch.relu(_lin(tgt, sd, p + "linear1.")), sd, p + "linear2.")
        tgt = _ln(tgt + ff, sd, p + "norm3.")
    return _lin(tgt, sd, prefix + "output_proj.")


def score_v1(sd: SD, tokens: Tensor, recon: Tensor) -> Tensor:
    """Shopformer.compute_normality_score, shopformer/models/shopformer.py:150-178:
    mean over (s, j) of (recon - (tokens + pe[:S]))^2, pe from the *facade's* buffer."""
    S = tokens.shape[1]
    pe = sd["pos_encoder.pe"].to(tokens.dtype)[:, :S]
    return ((recon - (tokens + pe)) ** 2).mean(dim=(1, 2))


# --------------------------------------------------------------------------- transformer v2
def reconstruct_v2(sd: SD, tokens: Tensor, nhead: int, *, prefix: str = "transformer.") -> Tensor:
    """ShopformerTransformer.forward, shopformer_2/models/transformer.py:147-194 with
    nn.TransformerEncoder/Decoder configured at :105-136 (norm_first, exact-erf GELU,
    final LayerNorm on both stacks, decoder target = encoder input, no masks)."""
    B, S, _ = tokens.shape
    x = tokens
    if (prefix + "input_projection.weight") in sd:
        x = _lin(x, sd, prefix + "input_projection.")
    x = x + sd[prefix + "pos_encoder.pe"].to(x.dtype)[:, :S]
    n_enc = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith(prefix + "encoder.layers."))
    n_dec = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith(prefix + "decoder.layers."))
    m = x
    for i in range(n_enc):
        p = f"{prefix}encoder.layers.{i}."
        h = _ln(m, sd, p + "norm1.")
        m = m + _mha(h, h, sd, p + "self_attn.", nhead)
        h = _ln(m, sd, p + "norm2.")
        m = m + _lin(F.gelu(_lin(h, sd, p + "linear1.")), sd, p + "linear2.")
    mem = _ln(m, sd, prefix + "encoder.norm.")
    y = x
    for i in range(n_dec):
        p = f"{prefix}decoder.layers.{i}."
        h = _ln(y, sd, p + "norm1.")
        y = y + _mha(h, h, sd, p + "self_attn.", nhead)
        h = _ln(y, sd, p + "norm2.")
        y = y + _mha(h, mem, sd, p + "multihead_attn.", nhead)
        h = _ln(y, sd, p + "norm3.")
        y = y + _lin(F.gelu(_lin(h, sd, p + "linear1.")), sd, p + "linear2.")
    y = _ln(y, sd, prefix + "decoder.norm.")
    if (prefix + "output_projection.weight") in sd:
        y = _lin(y, sd, prefix + "output_projection.")
    return y


def score_v2(tokens: Tensor, recon: Tensor, reduction: str = "mean") -> Tensor:
    """Shopformer.compute_anomaly_score, shopformer_2/models/shopformer.py:178-186."""
    if reduction == "mean":
        return ((tokens - recon) ** 2).mean(dim=(1, 2))
    if reduction == "none":
        return ((tokens - recon) ** 2).mean(dim=2)
    raise ValueError(f"Unknown reduction: {reduction}")


# --------------------------------------------------------------------------- whole path
def score_windows(sd: SD, poses: Tensor, *, variant: int, strides: Sequence[int], nhead: int,
                  pool_tokens: Optional[int] = None, reduction: str = "mean",
                  dtype: torch.dtype = torch.float32) -> Dict[str, Tensor]:
    """poses -> {tokens, recon, score}: the path of SURVEY 3.1 / 3.2 without the GCAE decoder."""
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    tok = tokenize(sd, poses.to(dtype), strides, pool_tokens=pool_tokens)
    if variant == 1:
        rec = reconstruct_v1(sd, tok, nhead)
        sc = score_v1(sd, tok, rec)
    else:
        rec = reconstruct_v2(sd, tok, nhead)
        sc = score_v2(tok, rec, reduction)
    return {"tokens": tok, "recon": rec, "score": sc}
