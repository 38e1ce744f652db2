"""Data loading for the variant-1 drop-in (reference: shopformer/data/__init__.py)."""
from .poselift_dataset import PoseLiftDataset, PoseLiftDataModule

__all__ = ["PoseLiftDataset", "PoseLiftDataModule"]
