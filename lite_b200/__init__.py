"""lite_b200 — B200-native pseudo-likelihood hot path of LiTe (kien-kieu/lite).

The product is the C-ABI shared library ``libliteb200.so`` (CUDA, sm_100a) and
the C++ facade in ``include/ttessel.h``.  This Python package is only a thin
ctypes binding used by the tests and ``bench.py``; it contains no compute path
of its own and raises if the native library is missing.
"""
from .capi import (LiteError, Context, TessSoA, FEATURE_IDS, lib_path, load_library,
                   nccl_unique_id, shard_range, pl_combine, SMF_COLS)

from .host import HostTess, generate_crtt

__all__ = ["HostTess", "generate_crtt", "LiteError", "Context", "TessSoA", "FEATURE_IDS", "lib_path", "load_library",
           "nccl_unique_id", "shard_range", "pl_combine", "SMF_COLS"]
