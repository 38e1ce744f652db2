"""Utilities of the variant-2 drop-in (reference: shopformer_2/utils/__init__.py)."""
from .device import get_device, check_mps_availability
from .config import load_config
from .metrics import compute_metrics, compute_auc_roc, compute_auc_pr

__all__ = ["get_device", "check_mps_availability", "load_config", "compute_metrics", "compute_auc_roc", "compute_auc_pr"]
