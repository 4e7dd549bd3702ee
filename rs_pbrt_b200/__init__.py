"""rs_pbrt_b200 -- B200-native PathIntegrator hot path for rs_pbrt behind a C ABI.

The product is rs_pbrt_b200/librs_pbrt_b200.so (hand-written CUDA for sm_100a + the C++ host mirror);
this package only holds the ctypes bindings and the synthetic scene generators used by tests/bench.
Importing the package does not load the library; the first call does, and fails loudly if it is missing.
"""
from . import _abi  # noqa: F401
from .host import GpuScene, HostScene, PbrtError, bvh_build, pin_description, render_multi, unpin_description  # noqa: F401

__all__ = ["GpuScene", "HostScene", "PbrtError", "bvh_build", "render_multi", "pin_description", "unpin_description"]
