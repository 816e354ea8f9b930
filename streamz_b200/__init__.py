"""streamz_b200: B200-native (sm_100a) implementation of StreamZ's data-parallel hot path.

The product is the CUDA shared library ``streamz_b200/lib/libstreamz_b200.so`` behind the C ABI of
``include/streamz_b200.h``; this package is its ctypes binding plus a Python mirror of the reference's Rust API
(``streamz_b200.api``).  Importing it without the built library raises -- there is no fallback path.
"""
from . import _native  # noqa: F401  (raises loudly when the library is missing)
from .api import *  # noqa: F401,F403
from .api import Context, FeatureExtractor, SimpleNeuralNet  # noqa: F401

__version__ = "0.1.0"
