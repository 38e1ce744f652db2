"""Model components of the variant-1 drop-in (reference: shopformer/models/__init__.py:12-24)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:          # makes `shopformer_b200` importable when a script runs from this directory
    sys.path.append(_ROOT)

from .gcae import GCAE, GraphConvolution, TemporalConvolution, STGCNBlock  # noqa: E402
from .transformer import ShopformerTransformer, PositionalEncoding  # noqa: E402
from .shopformer import Shopformer  # noqa: E402

__all__ = ["GCAE", "GraphConvolution", "TemporalConvolution", "STGCNBlock", "ShopformerTransformer",
           "PositionalEncoding", "Shopformer"]
