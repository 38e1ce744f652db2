"""Drop-in for the reference's ``shopformer_2/`` package (reference: shopformer_2/__init__.py)."""
__version__ = "2.0.0"
