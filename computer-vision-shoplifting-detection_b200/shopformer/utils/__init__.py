"""Host-side metrics (reference: shopformer/utils/__init__.py)."""
from .metrics import compute_auc_roc, compute_metrics

__all__ = ["compute_auc_roc", "compute_metrics"]
