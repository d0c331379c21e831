"""Host-side Python mirror of the reference's engine interface on top of libdsocr.so (ctypes).

The CUDA library is the product; this package only marshals host buffers into the C ABI declared in
include/dsocr.h.  There is no CPU fallback: importing `lib()` fails loudly if the shared library is missing.
"""
from .binding import lib, DsocrError, LIB_PATH  # noqa: F401
