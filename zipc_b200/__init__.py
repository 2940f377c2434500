"""zipc_b200 -- B200-native DEFLATE / zlib / ZIP hot path behind zipc's API.

The product is zipc_b200/libzipc_b200.so (hand-written sm_100a CUDA kernels behind the C ABI in
include/zipc_b200.h).  `zipc_deflate` and `zipc` mirror the reference's two modules on top of it.
There is no CPU fallback: compute calls fail loudly without the library or without a CUDA device.
"""
from . import _lib  # noqa: F401
from . import zipc_deflate  # noqa: F401
from . import synth  # noqa: F401
from . import zipc  # noqa: F401

__all__ = ["_lib", "zipc_deflate", "zipc", "synth"]
