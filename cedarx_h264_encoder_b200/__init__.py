"""B200-native drop-in for the encode path of libv/cedarx_h264_encoder.

The product is the C-ABI shared library `libcedar_b200.so` (include/cedar_b200.h) built from
csrc/ (hand-written sm_100a CUDA kernels + host C header writer).  This Python package is the thin
host-side mirror used by the tests and the benchmark: ctypes over that C ABI, nothing else.
There is no CPU fallback: importing works anywhere, but every encode call needs the built
library and a CUDA device and raises loudly otherwise.
"""
from .api import (CedarConfig, CedarIO, Encoder, Pipe, LibraryMissing, build_library, library_path, load_library,
                  write_pps, write_sps, slice_header_bits, FORMAT_NV12, FORMAT_NV16, ENTROPY_CAVLC, ENTROPY_CABAC)

__all__ = ["CedarConfig", "CedarIO", "Encoder", "Pipe", "LibraryMissing", "build_library", "library_path", "load_library",
           "write_pps", "write_sps", "slice_header_bits", "FORMAT_NV12", "FORMAT_NV16", "ENTROPY_CAVLC",
           "ENTROPY_CABAC"]
