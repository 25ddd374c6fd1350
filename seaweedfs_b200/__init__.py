"""seaweedfs_b200 — B200-native Reed–Solomon erasure coding behind SeaweedFS's
weed/storage/erasure_coding surface.

The product is the C-ABI shared library ``libswec.so`` (declared in ``include/swec.h``, built from
``seaweedfs_b200/csrc`` by ``seaweedfs_b200/build.py``).  This package is the thin ctypes binding
used by the tests and the benchmark; it has no CPU implementation and raises if the library or a
CUDA device is missing.
"""
from . import erasure_coding  # noqa: F401
from ._native import SwecError, lib, library_path  # noqa: F401

__all__ = ["erasure_coding", "SwecError", "lib", "library_path"]
