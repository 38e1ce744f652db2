"""Data package of the variant-2 drop-in (reference: shopformer_2/data/__init__.py)."""
from .poselift_dataset import PoseLiftDataset, PoseAugmentor

__all__ = ["PoseLiftDataset", "PoseAugmentor"]
