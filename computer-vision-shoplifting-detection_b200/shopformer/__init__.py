"""Drop-in for the reference's ``shopformer/`` package (same classes, constructor
signatures and ``state_dict`` keys); eval-mode inference runs on the sm_100a kernels of
``shopformer_b200``.  Reference: shopformer/__init__.py:28-31."""
from .models import Shopformer, GCAE, ShopformerTransformer

__version__ = "0.1.0"
__all__ = ["Shopformer", "GCAE", "ShopformerTransformer"]
