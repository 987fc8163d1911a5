"""B200-native Aho-Corasick / Meyer scan engine -- Python plumbing over the C-ABI of libac75.so.

The product is the shared library built from csrc/ (host C + hand-written sm_100a CUDA).  This package only binds its
C-ABI (include/aho_corasick.h, include/acm_b200.h) with ctypes so that tests and bench.py can drive it; it contains no
matching logic and no CPU fallback: every scan goes through acm_b200_scan_ex and fails loudly without the library or a GPU.
"""
from .binding import (  # noqa: F401
    MATCH_DTYPE,
    AcmError,
    Machine,
    build_library,
    device_count,
    lib,
    library_path,
)
from .shard import plan_shards, sharded_scan  # noqa: F401
