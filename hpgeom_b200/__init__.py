"""hpgeom_b200 -- B200 (sm_100a) implementation of hpgeom's elementwise HEALPix conversions.

``from hpgeom_b200 import *`` gives the same names as ``from hpgeom import *`` for the
conversion path (reference: hpgeom/__init__.py:12-14, hpgeom/hpgeom.py:21-54).
"""
from .hpgeom import *  # noqa: F401,F403
from .hpgeom import __all__ as _api_all
from . import healpy_compat  # noqa: F401
from . import sharding  # noqa: F401

__version__ = "0.1.0"
__all__ = list(_api_all) + ["healpy_compat", "sharding"]
