"""The library's DSQ writer (csrc/dsq_writer.cpp through the host-only dsocr_dsq_writer_* calls) against the reference's
writer tests, crates/dsq-writer/tests/writer.rs (:17-126, :186-277: same inputs, same assertions; the two
`*_bytes_match_candle_from_float` cases need candle and stay unpinned), against the committed gguf-py vectors, and against
the oracle's container writer byte for byte.  Files are read back with the library's own reader (dsocr_dsq_inspect) and
with the oracle's.  Runs without a GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import dsq
from tests.test_dsq_reader_cpu import Hdr, Rec, lib  # noqa: F401  (fixture + record structs)

Q8_0, Q4K, Q6K, F16, BF16, F32 = 8, 12, 14, 1, 16, 0
FP = C.POINTER(C.c_float)


def _writer(lib, path, default=Q8_0):
    lib.dsocr_dsq_writer_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.dsocr_dsq_writer_add_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, FP, FP]
    lib.dsocr_dsq_writer_add_quantized_bytes.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                         C.POINTER(C.c_uint8), C.c_size_t, FP]
    lib.dsocr_dsq_writer_finalize.argtypes = [C.c_void_p]
    lib.dsocr_dsq_writer_destroy.argtypes = [C.c_void_p]
    h = C.c_void_p()
    st = lib.dsocr_dsq_writer_create(str(path).encode(), b"candle-test", b"unit-test", b"CPU", default, C.byref(h))
    assert st == 0, lib.dsocr_last_error().decode()
    return h


def _add(lib, h, name, w, dt, bias=None):
    w = np.ascontiguousarray(w, dtype=np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    return lib.dsocr_dsq_writer_add_tensor(h, name.encode(), w.shape[0], w.shape[1], dt, w.ctypes.data_as(FP),
                                           b.ctypes.data_as(FP) if b is not None else None)


def _inspect(lib, path, cap=8):
    hdr, recs = Hdr(), (Rec * cap)()
    assert lib.dsocr_dsq_inspect(str(path).encode(), C.byref(hdr), recs, cap) == 0, lib.dsocr_last_error().decode()
    return hdr, recs


def test_writes_q8_tensor(lib, tmp_path):
    """writer.rs:17-60."""
    h = _writer(lib, tmp_path / "snapshot")
    w = (np.arange(2 * 32, dtype=np.float32) * 0.25 - 3.0).reshape(2, 32)
    bias = np.array([0.5, -0.25], np.float32)
    assert _add(lib, h, "linear.weight", w, Q8_0, bias) == 0
    assert lib.dsocr_dsq_writer_finalize(h) == 0
    path = tmp_path / "snapshot.dsq"  # output.with_extension("dsq")
    hdr, recs = _inspect(lib, path)
    assert hdr.tensor_count == 1 and hdr.candle_version == b"candle-test" and hdr.model_id == b"unit-test" and hdr.backend == b"CPU"
    r = recs[0]
    assert (r.name, r.out_dim, r.in_dim, r.q_dtype) == (b"linear.weight", 2, 32, Q8_0)
    assert r.q_len == 2 * (32 // 32) * 34 and r.bias_len == 8 and r.bias_dtype == 4
    raw = path.read_bytes()
    assert np.array_equal(np.frombuffer(raw[r.bias_offset:r.bias_offset + 8], np.float32), bias)
    payload = raw[r.q_offset:r.q_offset + r.q_len]
    assert any(payload) and payload == dsq.quantize_q8_0(w)


@pytest.mark.parametrize("dt,nb", [(Q4K, 144), (Q6K, 210)])
def test_writes_k_quant_tensors(lib, tmp_path, dt, nb):
    """writer.rs:62-126: record dims / dtype / byte length, payload not all zero; plus the dequantisation error of the
    blocks (the quantisers are ggml-style, not pinned to candle)."""
    h = _writer(lib, tmp_path / "snapshot_k", default=dt)
    out_dim, in_dim = 4, 512
    rng = np.random.RandomState(dt)
    w = (rng.randn(out_dim, in_dim) * 0.05).astype(np.float32)
    w[1, 256:512] = 0.0   # an all-zero super-block
    w[2, :32] = 0.3       # a constant sub-block
    assert _add(lib, h, "layer.weight", w, dt) == 0
    assert lib.dsocr_dsq_writer_finalize(h) == 0
    path = tmp_path / "snapshot_k.dsq"
    hdr, recs = _inspect(lib, path)
    r = recs[0]
    assert (r.out_dim, r.in_dim, r.q_dtype) == (out_dim, in_dim, dt) and r.q_len == out_dim * (in_dim // 256) * nb
    assert hdr.default_qdtype == dt and hdr.block_size == 256
    payload = path.read_bytes()[r.q_offset:r.q_offset + r.q_len]
    assert any(payload)
    deq = dsq.dequantize(payload, dt, out_dim, in_dim)
    assert np.isfinite(deq).all() and (deq[1, 256:] == 0).all()
    rel = np.sqrt(((deq - w) ** 2).mean()) / np.sqrt((w ** 2).mean())
    simple = dsq.dequantize(dsq._QUANT[dt](w), dt, out_dim, in_dim)
    rel_simple = np.sqrt(((simple - w) ** 2).mean()) / np.sqrt((w ** 2).mean())
    print(f"[dsq writer] dtype {dt}: relative rmse {rel:.4f} (oracle's min/max quantiser {rel_simple:.4f})")
    assert rel < (0.09 if dt == Q4K else 0.025) and rel <= rel_simple * 1.05


def test_writes_float_payloads(lib, tmp_path):
    """writer.rs:186-277: f32 payload is the input bytes; f16 / bf16 payloads are the values rounded to nearest-even."""
    import torch

    w = np.array([[0.5, -1.25, 2.0], [0.125, -0.75, 1.5]], np.float32)
    bias = np.array([0.25, -0.5], np.float32)
    h = _writer(lib, tmp_path / "snapshot_f32")
    assert _add(lib, h, "dense.weight", w, F32, bias) == 0
    wr = np.random.RandomState(3).randn(2, 3).astype(np.float32)
    assert _add(lib, h, "dense.f16", wr, F16) == 0
    assert _add(lib, h, "dense.bf16", wr, BF16) == 0
    assert lib.dsocr_dsq_writer_finalize(h) == 0
    path = tmp_path / "snapshot_f32.dsq"
    hdr, recs = _inspect(lib, path)
    raw = path.read_bytes()
    r = recs[0]
    assert r.q_dtype == F32 and r.q_len == 2 * 3 * 4 and raw[r.q_offset:r.q_offset + r.q_len] == w.tobytes() and r.bias_len == 8
    assert recs[1].q_dtype == F16 and recs[1].q_len == 12
    assert raw[recs[1].q_offset:recs[1].q_offset + 12] == wr.astype(np.float16).tobytes()
    assert recs[2].q_dtype == BF16 and raw[recs[2].q_offset:recs[2].q_offset + 12] == torch.from_numpy(wr).to(torch.bfloat16).view(torch.int16).numpy().tobytes()


def test_q8_quantiser_bit_exact_with_golden_gguf(lib, tmp_path):
    z = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "dsq_blocks.npz"))
    h = _writer(lib, tmp_path / "g")
    assert _add(lib, h, "w", z["q8_weight"], Q8_0) == 0
    assert lib.dsocr_dsq_writer_finalize(h) == 0
    _, recs = _inspect(lib, tmp_path / "g.dsq")
    raw = (tmp_path / "g.dsq").read_bytes()
    assert raw[recs[0].q_offset:recs[0].q_offset + recs[0].q_len] == z["q8_bytes"].tobytes()


def test_rejections_use_the_reference_wording(lib, tmp_path):
    h = _writer(lib, tmp_path / "bad")
    w = np.zeros((2, 32), np.float32)
    assert _add(lib, h, "a", w, Q8_0) == 0
    assert _add(lib, h, "a", w, Q8_0) != 0 and "already exists" in lib.dsocr_last_error().decode()            # DuplicateTensor
    assert _add(lib, h, "b", np.zeros((2, 30), np.float32), Q8_0) != 0 and "not divisible by block size 32" in lib.dsocr_last_error().decode()
    q = np.zeros(10, np.uint8)
    st = lib.dsocr_dsq_writer_add_quantized_bytes(h, b"c", 2, 256, Q4K, q.ctypes.data_as(C.POINTER(C.c_uint8)), 10, None)
    assert st != 0 and "expected 288 elements but received 10" in lib.dsocr_last_error().decode()              # DimensionMismatch
    st = lib.dsocr_dsq_writer_add_quantized_bytes(h, b"d", 2, 3, F32, q.ctypes.data_as(C.POINTER(C.c_uint8)), 10, None)
    assert st != 0 and "expects quantized dtype" in lib.dsocr_last_error().decode()
    lib.dsocr_dsq_writer_destroy(h)


def test_file_is_byte_identical_to_the_oracle_writer(lib, tmp_path):
    """Same tensors through the oracle's write_snapshot (crates/dsq layout restated in numpy) and through the library:
    header, records, offsets and payload must agree byte for byte (Q8_0 blocks and an externally quantised Q4_K tensor)."""
    rng = np.random.RandomState(11)
    w1 = (rng.randn(4, 64) * 0.1).astype(np.float32)
    w2 = (rng.randn(2, 256) * 0.1).astype(np.float32)
    bias = rng.randn(4).astype(np.float32)
    q2 = dsq.quantize_q4k(w2)
    ref = tmp_path / "ref.dsq"
    dsq.write_snapshot(str(ref), dsq.Q8_0, [("a.weight", 4, 64, dsq.Q8_0, dsq.quantize_q8_0(w1), bias.tobytes()),
                                            ("b.weight", 2, 256, dsq.Q4K, q2, None)],
                       model_id="unit-test", backend="CPU", candle_version="candle-test")
    h = _writer(lib, tmp_path / "mine")
    assert _add(lib, h, "a.weight", w1, Q8_0, bias) == 0
    qb = np.frombuffer(q2, np.uint8)
    assert lib.dsocr_dsq_writer_add_quantized_bytes(h, b"b.weight", 2, 256, Q4K, qb.ctypes.data_as(C.POINTER(C.c_uint8)), len(q2), None) == 0
    assert lib.dsocr_dsq_writer_finalize(h) == 0
    assert (tmp_path / "mine.dsq").read_bytes() == ref.read_bytes()


def test_writer_reader_roundtrip_random_tensor_sets(lib, tmp_path):
    """Seeded random snapshots (mixed dtypes, odd names, with / without bias): what the library writes, the library's
    reader and the oracle's reader both list, and the dequantised payload stays within the format's error of the input."""
    rng = np.random.RandomState(77)
    for case in range(6):
        h = _writer(lib, tmp_path / f"rt{case}", default=[Q8_0, Q4K, Q6K][case % 3])
        want = []
        for t in range(int(rng.randint(1, 6))):
            dt = [Q8_0, Q4K, Q6K, F32, F16, BF16][int(rng.randint(0, 6))]
            blk = {Q8_0: 32, Q4K: 256, Q6K: 256}.get(dt, 1)
            out_dim, in_dim = int(rng.randint(1, 9)) * 8, int(rng.randint(1, 4)) * blk * (1 if blk > 1 else 8)
            w = (rng.randn(out_dim, in_dim) * 10 ** rng.uniform(-3, 1)).astype(np.float32)
            bias = rng.randn(out_dim).astype(np.float32) if rng.rand() < 0.5 else None
            name = f"model.layers.{t}.weird name/{case}.weight"
            assert _add(lib, h, name, w, dt, bias) == 0, lib.dsocr_last_error().decode()
            want.append((name, dt, w, bias))
        assert lib.dsocr_dsq_writer_finalize(h) == 0
        path = tmp_path / f"rt{case}.dsq"
        hdr, recs = _inspect(lib, path, cap=8)
        assert hdr.tensor_count == len(want)
        _, orecs, data = dsq.read_snapshot(str(path))
        for i, (name, dt, w, bias) in enumerate(want):
            r, o = recs[i], orecs[name]
            assert r.name.decode() == name and (r.q_dtype, r.out_dim, r.in_dim) == (dt, w.shape[0], w.shape[1]) == (o.q_dtype, o.out_dim, o.in_dim)
            assert (r.bias_len != 0) == (bias is not None)
            if bias is not None:
                assert np.array_equal(np.frombuffer(data[r.bias_offset:r.bias_offset + r.bias_len], np.float32), bias)
            deq = dsq.dequantize(data[r.q_offset:r.q_offset + r.q_len], dt, w.shape[0], w.shape[1])
            tol = {Q8_0: 0.01, Q4K: 0.12, Q6K: 0.04, F32: 0.0, F16: 1e-3, BF16: 8e-3}[dt]
            err = np.abs(deq - w).max() / max(1e-30, np.abs(w).max())
            assert err <= tol, (name, dt, err)
