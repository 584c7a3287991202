"""Import shim for ``import datasets`` (datasets/__init__.py): ``datasets.__dict__["RainDrop"]``."""
from wavedm_b200.raindrop_data import RainDrop, RainDropDataset  # noqa: F401

__all__ = ["RainDrop"]
