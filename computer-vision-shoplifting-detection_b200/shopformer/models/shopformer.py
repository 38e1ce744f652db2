"""Variant-1 facade: the drop-in boundary for ``shopformer/{train,evaluate,inference}.py``.

Reference surface: shopformer/models/shopformer.py (``Shopformer`` :22-278, ``ShopformerStage1``
:281-321, ``ShopformerStage2`` :324-389).  Under ``model.eval()`` + ``torch.no_grad()`` on a CUDA
tensor, ``forward`` is ONE native call (``sf_score_windows``: tokenizer + transformer + fused
reconstruction-error score) and the GCAE pose decoder output is produced lazily, only if the caller
reads ``out['gcae_reconstructed']``.  In training the ATen composition below runs (autograd,
train-mode BatchNorm, dropout).
"""
from typing import Any, Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from shopformer_b200.engine import EngineConfig
from shopformer_b200.facade import EngineCacheMixin, LazyOutput
from shopformer_b200.modules import adopt, wants_native
from shopformer_b200.native import SF_VARIANT_SHOPFORMER
from shopformer_b200.ops import model_handle

from .gcae import GCAE, GCAEEncoder  # noqa: F401
from .transformer import PositionalEncoding, ShopformerTransformer

__all__ = ["Shopformer", "ShopformerStage1", "ShopformerStage2"]


class Shopformer(nn.Module, EngineCacheMixin):
    def __init__(self, in_channels: int = 2, hidden_channels: int = 64, latent_channels: int = 8,
                 num_keypoints: int = 17, seq_len: int = 12, num_tokens: int = 2, gcae_layers: int = 4,
                 transformer_heads: int = 2, transformer_layers: int = 2, transformer_ff_dim: int = 64,
                 dropout: float = 0.1, layout: str = "coco", freeze_tokenizer: bool = False):
        super().__init__()
        self.in_channels, self.num_keypoints = in_channels, num_keypoints
        self.seq_len, self.num_tokens, self.latent_channels = seq_len, num_tokens, latent_channels
        self.embedding_dim = latent_channels * num_keypoints
        self.gcae = GCAE(in_channels=in_channels, hidden_channels=hidden_channels, latent_channels=latent_channels,
                         num_keypoints=num_keypoints, seq_len=seq_len, num_tokens=num_tokens, num_layers=gcae_layers,
                         dropout=dropout, layout=layout)
        self.transformer = ShopformerTransformer(d_model=self.embedding_dim, nhead=transformer_heads,
                                                 num_encoder_layers=transformer_layers,
                                                 num_decoder_layers=transformer_layers,
                                                 dim_feedforward=transformer_ff_dim, dropout=dropout)
        self.pos_encoder = PositionalEncoding(self.embedding_dim, max_len=100, dropout=0.0)
        self._sf_meta = dict(heads=transformer_heads, layers=transformer_layers, ff=transformer_ff_dim)
        adopt(self, self.gcae.encoder, self.transformer)
        self.freeze_tokenizer = freeze_tokenizer
        if freeze_tokenizer:
            self._freeze_gcae_encoder()

    # -- native model description ------------------------------------------------------------
    def _sf_config(self) -> EngineConfig:
        enc = self.gcae.encoder
        return EngineConfig(variant=SF_VARIANT_SHOPFORMER, in_channels=self.in_channels,
                            num_keypoints=self.num_keypoints, channels=list(enc._channels), strides=list(enc.strides),
                            d_model=self.embedding_dim, n_heads=self._sf_meta["heads"],
                            n_enc_layers=self._sf_meta["layers"], n_dec_layers=self._sf_meta["layers"],
                            d_ff=self._sf_meta["ff"], pool_tokens=0)

    # -- reference API -------------------------------------------------------------------------
    def _freeze_gcae_encoder(self):
        for p in self.gcae.encoder.parameters():
            p.requires_grad = False

    def unfreeze_tokenizer(self):
        for p in self.gcae.encoder.parameters():
            p.requires_grad = True
        self.freeze_tokenizer = False

    def tokenize(self, poses: torch.Tensor) -> torch.Tensor:
        return self.gcae.encode(poses)

    def reconstruct_tokens(self, tokens: torch.Tensor) -> torch.Tensor:
        return self.transformer(tokens)

    def compute_normality_score(self, tokens: torch.Tensor, reconstructed: torch.Tensor) -> torch.Tensor:
        if wants_native(self, tokens):
            return torch.ops.shopformer_b200.normality_score(tokens, reconstructed, model_handle(self._sf_engine()), "mean")
        target = tokens + self.pos_encoder.pe[:, :tokens.size(1), :].expand(tokens.size(0), -1, -1)
        return F.mse_loss(reconstructed, target, reduction="none").mean(dim=[1, 2])

    def forward(self, poses: torch.Tensor, return_tokens: bool = False) -> Dict[str, torch.Tensor]:
        if wants_native(self, poses):
            x = self.gcae.encoder._as_bctv(poses)
            score, tokens, recon = torch.ops.shopformer_b200.score_fused_full(x, model_handle(self._sf_engine()),
                                                                              self._sf_resolve_precision())
            out = LazyOutput({"normality_score": score, "reconstructed_tokens": recon},
                             {"gcae_reconstructed": lambda: self._decode_eval(tokens)})
            if return_tokens:
                out["tokens"] = tokens
            return out
        tokens = self.tokenize(poses)
        recon = self.reconstruct_tokens(tokens)
        out = {"normality_score": self.compute_normality_score(tokens, recon), "reconstructed_tokens": recon,
               "gcae_reconstructed": self.gcae.decode(tokens)}
        if return_tokens:
            out["tokens"] = tokens
        return out

    def _decode_eval(self, tokens: torch.Tensor) -> torch.Tensor:
        """The pose decoder with the semantics of the call that produced `tokens` (eval-mode BatchNorm, no dropout, no
        autograd graph), whatever mode the model is in when the lazy output is finally read."""
        dec = self.gcae.decoder
        was_training = dec.training
        try:
            dec.eval()
            with torch.no_grad():
                return dec(tokens)
        finally:
            dec.train(was_training)

    def predict(self, poses: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
        with torch.no_grad():
            return (self.forward(poses)["normality_score"] > threshold).long()

    def get_anomaly_scores(self, poses: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return self.forward(poses)["normality_score"]

    @classmethod
    def from_config(cls, config: Dict[str, Any]) -> "Shopformer":
        defaults = dict(in_channels=2, hidden_channels=64, latent_channels=8, num_keypoints=17, seq_len=12,
                        num_tokens=2, gcae_layers=4, transformer_heads=2, transformer_layers=2,
                        transformer_ff_dim=64, dropout=0.1, layout="coco", freeze_tokenizer=False)
        return cls(**{k: config.get(k, v) for k, v in defaults.items()})


class ShopformerStage1(nn.Module):
    """Stage 1 wrapper: train the GCAE by pose reconstruction."""

    def __init__(self, shopformer: Shopformer):
        super().__init__()
        self.gcae = shopformer.gcae

    def forward(self, poses: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.gcae(poses)

    def compute_loss(self, poses: torch.Tensor, reconstructed: torch.Tensor) -> torch.Tensor:
        return F.mse_loss(reconstructed, poses)


class ShopformerStage2(nn.Module):
    """Stage 2 wrapper: frozen tokenizer, train the transformer on token reconstruction."""

    def __init__(self, shopformer: Shopformer):
        super().__init__()
        self.shopformer = shopformer
        for p in self.shopformer.gcae.encoder.parameters():
            p.requires_grad = False

    def forward(self, poses: torch.Tensor) -> Dict[str, torch.Tensor]:
        with torch.no_grad():
            tokens = self.shopformer.tokenize(poses)
        recon = self.shopformer.reconstruct_tokens(tokens)
        return {"tokens": tokens, "reconstructed_tokens": recon,
                "normality_score": self.shopformer.compute_normality_score(tokens, recon)}

    def compute_loss(self, tokens: torch.Tensor, reconstructed: torch.Tensor) -> torch.Tensor:
        pe = self.shopformer.pos_encoder.pe[:, :tokens.size(1), :].expand(tokens.size(0), -1, -1)
        return F.mse_loss(reconstructed, tokens + pe)
