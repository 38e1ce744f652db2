"""Variant-1 transformer (post-LN, ReLU, zero start token + shifted target, no masks).

Reference surface: shopformer/models/transformer.py (``PositionalEncoding`` :14-57,
``TransformerEncoderLayer`` :60-118, ``TransformerDecoderLayer`` :121-196,
``ShopformerTransformer`` :199-349).  Eval-mode CUDA inference is one native call
(``sf_reconstruct_tokens``); the bodies below are the autograd path used in training.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from shopformer_b200.ops import model_handle
from shopformer_b200.modules import (PostLNDecoderLayer as TransformerDecoderLayer,
                                     PostLNEncoderLayer as TransformerEncoderLayer, _Owned, sinusoid_table,
                                     wants_native)

__all__ = ["PositionalEncoding", "TransformerEncoderLayer", "TransformerDecoderLayer", "ShopformerTransformer"]


class PositionalEncoding(nn.Module):
    """x + pe[:, :S] (+ dropout in training); ``pe`` is a (1, max_len, d_model) buffer."""

    def __init__(self, d_model: int, max_len: int = 5000, dropout: float = 0.1):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.register_buffer("pe", sinusoid_table(d_model, max_len))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.dropout(x + self.pe[:, :x.size(1), :])


class ShopformerTransformer(nn.Module, _Owned):
    def __init__(self, d_model: int = 144, nhead: int = 2, num_encoder_layers: int = 2,
                 num_decoder_layers: int = 2, dim_feedforward: int = 64, dropout: float = 0.1,
                 max_seq_len: int = 100):
        super().__init__()
        self.d_model, self.nhead = d_model, nhead
        self.pos_encoder = PositionalEncoding(d_model, max_seq_len, dropout)
        self.encoder_layers = nn.ModuleList(
            [TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout) for _ in range(num_encoder_layers)])
        self.decoder_layers = nn.ModuleList(
            [TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout) for _ in range(num_decoder_layers)])
        self.output_proj = nn.Linear(d_model, d_model)
        self._init_weights()

    def _init_weights(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def encode(self, src: torch.Tensor) -> torch.Tensor:
        src = self.pos_encoder(src)
        for layer in self.encoder_layers:
            src = layer(src)
        return src

    def decode(self, tgt: torch.Tensor, memory: torch.Tensor) -> torch.Tensor:
        tgt = self.pos_encoder(tgt)
        for layer in self.decoder_layers:
            tgt = layer(tgt, memory)
        return tgt

    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        if wants_native(self, tokens):
            eng = self._engine()
            if eng is not None:
                return torch.ops.shopformer_b200.reconstruct_tokens(tokens, model_handle(eng), self._precision())
        memory = self.encode(tokens)
        start = torch.zeros(tokens.size(0), 1, self.d_model, device=tokens.device)
        shifted = torch.cat([start, tokens[:, :-1, :]], dim=1)
        return self.output_proj(self.decode(shifted, memory))

    def compute_reconstruction_error(self, tokens: torch.Tensor, reconstructed: torch.Tensor) -> torch.Tensor:
        return F.mse_loss(reconstructed, tokens, reduction="none").mean(dim=[1, 2])
