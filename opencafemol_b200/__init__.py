"""opencafemol_b200 — B200-native engine for OpenCafeMol's per-step coarse-grained MD hot path.

The product is the CUDA library ``opencafemol_b200/lib/libocm_b200.so`` (C ABI: include/ocm_b200.h);
this package is its Python front end (ctypes binding, TOML reader, workload generators).
"""
from .capi import Context, Library, OcmError, PairTopology, System, product_library  # noqa: F401

__version__ = "0.1.0"
