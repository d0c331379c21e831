"""ctypes loader for libdsocr.so."""
from __future__ import annotations

import ctypes
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent.parent / "lib" / "libdsocr.so"
_lib = None


class DsocrError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise DsocrError(f"{LIB_PATH} is missing: build it with `python deepseek-ocr.rs_b200/build.py` "
                             "(there is no CPU fallback)")
        _lib = ctypes.CDLL(str(LIB_PATH))
        _lib.dsocr_last_error.restype = ctypes.c_char_p
        _lib.dsocr_version.restype = ctypes.c_char_p
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().dsocr_last_error().decode("utf-8", "replace")
        raise DsocrError(f"{what} failed ({status}): {msg}")
