"""clann_b200 — B200-native (sm_100a CUDA) implementation of CLANN's build + search hot path.

The product is `clann_b200/lib/libclann_b200.so` (hand-written CUDA kernels + the C ABI of include/clann_b200.h).
This package is the host-side mirror of the reference's Rust API over that ABI (see clann_b200/api.py).
"""
from .api import *  # noqa: F401,F403
from .api import __all__  # noqa: F401
