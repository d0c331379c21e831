"""dsq_dequant64 (csrc/dsq_dequant.h): the per-(row, 64-wide k-block) dequantisation the dequant-fused tensor-core GEMM's
producer threads run, compiled for the host and compared bit for bit with the oracle's dequantisers (themselves pinned to
gguf-py) on random valid blocks and on the committed gguf-py vectors.  No GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import dsq
from tests.test_dsq_reader_cpu import lib  # noqa: F401

G = os.path.join(os.path.dirname(__file__), "golden", "dsq_blocks.npz")


def _deq(lib, dt, raw, rows, K):
    out = np.empty((rows, K), np.float32)
    buf = np.ascontiguousarray(np.frombuffer(raw, np.uint8))
    st = lib.dsocr_test_dsq_dequant64(dt, buf.ctypes.data_as(C.POINTER(C.c_uint8)), rows, K, out.ctypes.data_as(C.POINTER(C.c_float)))
    assert st == 0, lib.dsocr_last_error().decode()
    return out


@pytest.mark.parametrize("dt", [dsq.Q8_0, dsq.Q4K, dsq.Q6K])
def test_dequant64_matches_oracle_on_quantised_weights(lib, dt):
    rng = np.random.RandomState(dt)
    rows, K = 6, 1280 if dt != dsq.Q8_0 else 896
    w = (rng.randn(rows, K) * 0.05).astype(np.float32)
    raw = dsq._QUANT[dt](w)
    assert np.array_equal(_deq(lib, dt, raw, rows, K), dsq.dequantize(raw, dt, rows, K))


def test_dequant64_matches_golden_gguf_on_random_blocks(lib):
    z = np.load(G)
    for name, dt in (("q8_0", dsq.Q8_0), ("q4k", dsq.Q4K), ("q6k", dsq.Q6K)):
        raw, ref = z[f"{name}_raw"], z[f"{name}_deq"]
        K = ref.shape[1]
        if K % 64:
            continue
        got = _deq(lib, dt, raw.tobytes(), raw.shape[0], K)
        assert np.array_equal(got, ref) or np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max(), name
