"""Variant-2 transformer: pre-LN ``nn.TransformerEncoder/Decoder`` containers (GELU, final norms,
optional input/output projections when ``input_dim != d_model``).

Reference surface: shopformer_2/models/transformer.py (``PositionalEncoding`` :18-56,
``ShopformerTransformer`` :59-224, ``TransformerConfig`` :227-262, ``build_transformer`` :265-276).
The torch containers hold the parameters (identical ``state_dict`` keys) and serve training;
eval-mode CUDA inference is ``sf_reconstruct_tokens``.
"""
from typing import Optional

import torch
import torch.nn as nn

from shopformer_b200.modules import _Owned, sinusoid_table, wants_native
from shopformer_b200.ops import model_handle

__all__ = ["PositionalEncoding", "ShopformerTransformer", "TransformerConfig", "build_transformer"]


class PositionalEncoding(nn.Module):
    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 100):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.register_buffer("pe", sinusoid_table(d_model, max_len))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.dropout(x + self.pe[:, :x.size(1), :])


class ShopformerTransformer(nn.Module, _Owned):
    def __init__(self, input_dim: int = 144, d_model: int = 144, nhead: int = 12, num_encoder_layers: int = 4,
                 num_decoder_layers: int = 4, dim_feedforward: int = 64, dropout: float = 0.1,
                 max_seq_len: int = 100, activation: str = "gelu"):
        super().__init__()
        self.input_dim, self.d_model, self.nhead = input_dim, d_model, nhead
        self.dim_feedforward = dim_feedforward
        self.num_encoder_layers, self.num_decoder_layers = num_encoder_layers, num_decoder_layers
        self.activation_name = activation
        self.needs_projection = input_dim != d_model
        self.input_projection = nn.Linear(input_dim, d_model) if self.needs_projection else nn.Identity()
        self.output_projection = nn.Linear(d_model, input_dim) if self.needs_projection else nn.Identity()
        self.pos_encoder = PositionalEncoding(d_model, dropout, max_seq_len)
        layer_kw = dict(d_model=d_model, nhead=nhead, dim_feedforward=dim_feedforward, dropout=dropout,
                        activation=activation, batch_first=True, norm_first=True)
        self.encoder = nn.TransformerEncoder(nn.TransformerEncoderLayer(**layer_kw), num_layers=num_encoder_layers,
                                             norm=nn.LayerNorm(d_model))
        self.decoder = nn.TransformerDecoder(nn.TransformerDecoderLayer(**layer_kw), num_layers=num_decoder_layers,
                                             norm=nn.LayerNorm(d_model))
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def _embed(self, tokens: torch.Tensor) -> torch.Tensor:
        return self.pos_encoder(self.input_projection(tokens))

    def forward(self, tokens: torch.Tensor, src_mask: Optional[torch.Tensor] = None,
                tgt_mask: Optional[torch.Tensor] = None, src_key_padding_mask: Optional[torch.Tensor] = None,
                tgt_key_padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        unmasked = src_mask is None and tgt_mask is None and src_key_padding_mask is None and tgt_key_padding_mask is None
        if unmasked and self.activation_name == "gelu" and wants_native(self, tokens):
            eng = self._engine()
            if eng is not None:
                return torch.ops.shopformer_b200.reconstruct_tokens(tokens, model_handle(eng), self._precision())
        x = self._embed(tokens)
        memory = self.encoder(x, mask=src_mask, src_key_padding_mask=src_key_padding_mask)
        out = self.decoder(x, memory, tgt_mask=tgt_mask, memory_mask=src_mask,
                           tgt_key_padding_mask=tgt_key_padding_mask, memory_key_padding_mask=src_key_padding_mask)
        return self.output_projection(out)

    def encode(self, tokens: torch.Tensor) -> torch.Tensor:
        return self.encoder(self._embed(tokens))

    def decode(self, memory: torch.Tensor, tokens: torch.Tensor) -> torch.Tensor:
        return self.output_projection(self.decoder(self._embed(tokens), memory))


class TransformerConfig:
    """Defaults + config-dict adapter (``model.transformer`` section)."""
    INPUT_DIM = 144
    D_MODEL = 144
    NHEAD = 12
    NUM_ENCODER_LAYERS = 4
    NUM_DECODER_LAYERS = 4
    DIM_FEEDFORWARD = 64
    DROPOUT = 0.1
    MAX_SEQ_LEN = 100
    ACTIVATION = "gelu"

    @classmethod
    def from_config(cls, config: dict) -> dict:
        t = config.get("model", {}).get("transformer", {})
        return {
            "input_dim": t.get("input_dim", cls.INPUT_DIM),
            "d_model": t.get("d_model", cls.D_MODEL),
            "nhead": t.get("num_heads", cls.NHEAD),
            "num_encoder_layers": t.get("num_layers", cls.NUM_ENCODER_LAYERS),
            "num_decoder_layers": t.get("num_layers", cls.NUM_DECODER_LAYERS),
            "dim_feedforward": t.get("dim_feedforward", cls.DIM_FEEDFORWARD),
            "dropout": t.get("dropout", cls.DROPOUT),
        }


def build_transformer(config: dict) -> ShopformerTransformer:
    return ShopformerTransformer(**TransformerConfig.from_config(config))
