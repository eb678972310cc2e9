"""oss-b200: a B200-native engine for OpenSkyStacker's per-frame pipeline.

The product is `csrc/liboss_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/oss_b200.h) plus the C++ `ImageStacker` façade in csrc/.  This Python package
is only the ctypes door used by the tests and bench.py; it has no compute path of its
own and raises if the CUDA library is missing.
"""
from .binding import Engine, OssError, build, lib, lib_path  # noqa: F401
