"""Top-level stand-in for the reference's spatial_transformer.py: with this directory on sys.path,
`from spatial_transformer import *` (train_bundle_nobm.py:16) brings in this library's transformer / interpolate."""
import spatial_transformer3 as _st3  # noqa: F401  (loads the package alias)

from dovs_b200.spatial_transformer import interpolate, transformer  # noqa: E402,F401

__all__ = ['transformer', 'interpolate']
