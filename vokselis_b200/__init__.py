"""vokselis_b200 — B200-native (sm_100a CUDA) replacement for the raycast path of pudnax/vokselis.

The package holds only what that path needs: `csrc/` (CUDA kernels + the C ABI of
include/vokselis_rt.h), `host/` (C++ mirror of the reference's Rust host interface) and this thin
ctypes layer. Importing the package does not load the CUDA library; `vokselis_b200.rt.lib()` does and
raises if it is missing — there is no CPU fallback.
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
