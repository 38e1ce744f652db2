"""Model components of the variant-2 drop-in (reference: shopformer_2/models/__init__.py:3-14)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)

from .gcae import GCAE, GCAEEncoder, GCAEDecoder  # noqa: E402
from .transformer import ShopformerTransformer, PositionalEncoding  # noqa: E402
from .shopformer import Shopformer  # noqa: E402

__all__ = ["GCAE", "GCAEEncoder", "GCAEDecoder", "ShopformerTransformer", "PositionalEncoding", "Shopformer"]
