"""cubez_b200 — B200-native (sm_100a) implementation of the cubez per-step rigid-body pipeline.

The product is the C-ABI library `lib/libcubezcuda.so` (include/cubezcuda.h): hand-written CUDA
kernels for integrate+derive, narrowphase, the worst-first contact resolver, the fused
small-world step and the sort-based broadphase.  This package is the thin host side: ctypes
bindings, the reference's object API under its own names, the batched-world handle and the
synthetic scene builders.  There is no CPU fallback: importing `api` objects needs the built
library, and every compute call needs a CUDA device.
"""
from . import _abi, hostmath, scenes  # noqa: F401

__all__ = ["_abi", "hostmath", "scenes", "api"]


def __getattr__(name):
    if name == "api":
        import importlib
        return importlib.import_module(".api", __name__)
    raise AttributeError(name)
