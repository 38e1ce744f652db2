"""Variant-2 facade: the drop-in boundary for ``shopformer_2/{train,evaluate}.py``.

Reference surface: shopformer_2/models/shopformer.py (``Shopformer`` :20-293, ``build_shopformer``
:296-306).  ``compute_anomaly_score`` -- the call both evaluation loops make per batch
(shopformer_2/train.py:237-263, shopformer_2/evaluate.py:36-118) -- is ONE native call
(``sf_score_windows``) on CUDA tensors; the reference additionally runs and discards the GCAE pose
decoder there, which the native path does not compute at all.
"""
from typing import Any, Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from shopformer_b200.engine import EngineConfig
from shopformer_b200.facade import EngineCacheMixin
from shopformer_b200.modules import adopt, composite_eval_allowed
from shopformer_b200.native import SF_VARIANT_SHOPFORMER_2
from shopformer_b200.ops import model_handle

from .gcae import GCAE, GCAEEncoder  # noqa: F401
from .transformer import ShopformerTransformer, build_transformer  # noqa: F401

__all__ = ["Shopformer", "build_shopformer"]


class Shopformer(nn.Module, EngineCacheMixin):
    def __init__(self, config: Dict[str, Any]):
        super().__init__()
        self.config = config
        mc = config["model"]
        g = mc["gcae"]
        self.gcae = GCAE(in_channels=mc["in_channels"], hidden_channels=g["hidden_channels"],
                         latent_channels=g["latent_channels"], num_keypoints=mc["num_keypoints"],
                         seq_len=mc["seq_len"], num_tokens=mc["num_tokens"], num_layers=g.get("num_layers", 4),
                         dropout=g.get("dropout", 0.1))
        self.transformer = build_transformer(config)
        self._gcae_frozen = False
        self.num_keypoints, self.seq_len = mc["num_keypoints"], mc["seq_len"]
        self.num_tokens, self.latent_channels = mc["num_tokens"], g["latent_channels"]
        adopt(self, self.gcae.encoder, self.transformer)

    # -- native model description ------------------------------------------------------------
    def _sf_config(self) -> EngineConfig:
        enc, tr = self.gcae.encoder, self.transformer
        return EngineConfig(variant=SF_VARIANT_SHOPFORMER_2, in_channels=enc.in_channels,
                            num_keypoints=self.num_keypoints, channels=list(enc._channels), strides=list(enc.strides),
                            d_model=tr.d_model, n_heads=tr.nhead, n_enc_layers=tr.num_encoder_layers,
                            n_dec_layers=tr.num_decoder_layers, d_ff=tr.dim_feedforward,
                            pool_tokens=self.num_tokens if enc._needs_pooling else 0)

    # -- reference API -------------------------------------------------------------------------
    def freeze_gcae(self):
        for p in self.gcae.parameters():
            p.requires_grad = False
        self.gcae.eval()
        self._gcae_frozen = True

    def unfreeze_gcae(self):
        for p in self.gcae.parameters():
            p.requires_grad = True
        self._gcae_frozen = False

    def train(self, mode: bool = True):
        super().train(mode)
        if self._gcae_frozen and mode:
            self.gcae.eval()
        return self

    def _gcae_no_grad(self, poses: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        with torch.no_grad():
            recon, tokens = self.gcae(poses)
        return recon, tokens.detach()

    def forward(self, poses: torch.Tensor, return_all: bool = False):
        recon_poses, tokens = self._gcae_no_grad(poses) if self._gcae_frozen else self.gcae(poses)
        recon_tokens = self.transformer(tokens)
        return (recon_poses, tokens, recon_tokens) if return_all else recon_tokens

    def encode(self, poses: torch.Tensor) -> torch.Tensor:
        if self._gcae_frozen:
            with torch.no_grad():
                return self.gcae.encode(poses).detach()
        return self.gcae.encode(poses)

    def compute_anomaly_score(self, poses: torch.Tensor, reduction: str = "mean") -> torch.Tensor:
        if reduction not in ("mean", "none"):
            raise ValueError(f"Unknown reduction: {reduction}")
        self.eval()
        with torch.no_grad():
            if poses.is_cuda:
                x = self.gcae.encoder._as_bctv(poses)
                return torch.ops.shopformer_b200.score_fused(x, model_handle(self._sf_engine()), reduction, self._sf_resolve_precision())
            if not composite_eval_allowed():
                raise RuntimeError("shopformer_b200: compute_anomaly_score runs on CUDA (sm_100a) only; "
                                   "there is no CPU fallback")
            tokens = self.gcae.encode(poses)
            err = (tokens - self.transformer(tokens)) ** 2
            return err.mean(dim=(1, 2)) if reduction == "mean" else err.mean(dim=2)

    def compute_gcae_loss(self, poses: torch.Tensor) -> torch.Tensor:
        recon, _ = self.gcae(poses)
        return F.mse_loss(recon, poses)

    def compute_transformer_loss(self, poses: torch.Tensor) -> torch.Tensor:
        tokens = self.encode(poses)
        return F.mse_loss(self.transformer(tokens), tokens)

    def get_num_parameters(self, trainable_only: bool = True) -> Dict[str, int]:
        def count(mod):
            return sum(p.numel() for p in mod.parameters() if p.requires_grad or not trainable_only)
        return {"gcae": count(self.gcae), "transformer": count(self.transformer), "total": count(self)}

    @staticmethod
    def _sub_state(ckpt: Dict[str, Any], own_key: str, prefix: str) -> Dict[str, torch.Tensor]:
        if own_key in ckpt:
            return ckpt[own_key]
        if "model_state_dict" in ckpt:
            return {k[len(prefix):]: v for k, v in ckpt["model_state_dict"].items() if k.startswith(prefix)}
        return ckpt

    def load_gcae_checkpoint(self, checkpoint_path: str, strict: bool = True):
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        self.gcae.load_state_dict(self._sub_state(ckpt, "gcae_state_dict", "gcae."), strict=strict)

    def load_transformer_checkpoint(self, checkpoint_path: str, strict: bool = True):
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        self.transformer.load_state_dict(self._sub_state(ckpt, "transformer_state_dict", "transformer."), strict=strict)


def build_shopformer(config: Dict[str, Any]) -> Shopformer:
    return Shopformer(config)
