"""Top-level stand-in for the reference's spatial_transformer3.py: put this directory on sys.path ahead of the reference's
own files and `from spatial_transformer3 import transformer` (s_net_bundle_nobm.py:16) resolves here, unchanged."""
import importlib.util
import os
import sys

if 'dovs_b200' not in sys.modules:
    _pkg = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _spec = importlib.util.spec_from_file_location('dovs_b200', os.path.join(_pkg, '__init__.py'), submodule_search_locations=[_pkg])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules['dovs_b200'] = _mod
    _spec.loader.exec_module(_mod)

from dovs_b200.spatial_transformer3 import NAMED, interpolate, named_outputs, transformer  # noqa: E402,F401

__all__ = ['transformer', 'interpolate', 'named_outputs', 'NAMED']
