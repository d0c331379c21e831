"""The library's DSQ container reader (csrc/dsq.cpp, reached through the host-only C-ABI call dsocr_dsq_inspect) against
the reference's own reader tests, crates/dsq/tests/reader.rs: the same snapshot bytes (build_snapshot_bytes, :26-62), the
same six cases (:68-258) and the same substrings in the error messages.  Runs without a GPU."""
import ctypes as C
import struct

import numpy as np
import pytest

Q8_0, Q4K, Q6K, F16, BF16, F32 = 8, 12, 14, 1, 16, 0
BIAS_F32 = 4


class Rec(C.Structure):
    _fields_ = [("name", C.c_char * 192), ("out_dim", C.c_uint32), ("in_dim", C.c_uint32), ("q_dtype", C.c_uint32),
                ("q_offset", C.c_uint64), ("q_len", C.c_uint64), ("bias_offset", C.c_uint64), ("bias_len", C.c_uint64),
                ("bias_dtype", C.c_uint32), ("first_q_byte", C.c_uint8)]


class Hdr(C.Structure):
    _fields_ = [("version", C.c_uint32), ("default_qdtype", C.c_uint32), ("block_size", C.c_uint32), ("tensor_count", C.c_uint32),
                ("candle_version", C.c_char * 64), ("model_id", C.c_char * 128), ("backend", C.c_char * 32)]


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    from dsocr.binding import LIB_PATH, lib as load

    if not LIB_PATH.exists():
        ge.build()
    return load()


def _wstr(s: str) -> bytes:
    return struct.pack("<I", len(s)) + s.encode()


def build_snapshot_bytes(header_dtype, record_dtype, block_size, name, out_dim, in_dim, q_bytes, bias_bytes=None, version=1,
                         magic=b"DSQSNAP"):
    """reader.rs:26-62, byte for byte."""
    f = bytearray(magic) + struct.pack("<I", version) + _wstr("candle-test") + _wstr("model-id") + _wstr("CPU")
    f += struct.pack("<III", header_dtype, block_size, 1)
    record_size = (4 + len(name)) + 4 * 3 + 8 * 4 + 4
    q_offset = len(f) + record_size
    bias_offset = q_offset + len(q_bytes)
    f += _wstr(name) + struct.pack("<III", out_dim, in_dim, record_dtype) + struct.pack("<QQ", q_offset, len(q_bytes))
    if bias_bytes is not None:
        f += struct.pack("<QQI", bias_offset, len(bias_bytes), BIAS_F32)
    else:
        f += struct.pack("<QQI", 0, 0, 0)
    f += q_bytes
    if bias_bytes is not None:
        f += bias_bytes
    return bytes(f)


def inspect(lib, tmp_path, data, cap=4):
    p = tmp_path / "t.dsq"
    p.write_bytes(data)
    hdr, recs = Hdr(), (Rec * cap)()
    st = lib.dsocr_dsq_inspect(str(p).encode(), C.byref(hdr), recs, cap)
    msg = lib.dsocr_last_error().decode() if st != 0 else ""
    return st, msg, hdr, recs


def test_parses_valid_snapshot(lib, tmp_path):
    out_dim, in_dim = 64, 96
    q = bytes([0xAB]) * (out_dim * (in_dim // 32) * 34)
    bias = bytes(out_dim * 4)
    st, msg, hdr, recs = inspect(lib, tmp_path, build_snapshot_bytes(Q8_0, Q8_0, 32, "layer.q_proj.weight", out_dim, in_dim, q, bias))
    assert st == 0, msg
    assert hdr.tensor_count == 1 and hdr.default_qdtype == Q8_0 and hdr.block_size == 32
    assert (hdr.candle_version, hdr.model_id, hdr.backend) == (b"candle-test", b"model-id", b"CPU")
    r = recs[0]
    assert r.name == b"layer.q_proj.weight" and (r.out_dim, r.in_dim, r.q_dtype) == (out_dim, in_dim, Q8_0)
    assert r.q_len == len(q) and r.first_q_byte == 0xAB
    assert r.bias_len == len(bias) and r.bias_dtype == BIAS_F32 and r.bias_offset == r.q_offset + r.q_len


def test_rejects_unaligned_q8(lib, tmp_path):
    out_dim, in_dim = 64, 30
    q = bytes([0xCD]) * (out_dim * ((in_dim + 31) // 32) * 34)
    st, msg, _, _ = inspect(lib, tmp_path, build_snapshot_bytes(Q8_0, Q8_0, 32, "bad", out_dim, in_dim, q))
    assert st != 0 and "snapshot validation failed" in msg and "not divisible" in msg


@pytest.mark.parametrize("dt,fill,name", [(Q4K, 0xEE, "layer.o_proj.weight"), (Q6K, 0xAA, "layer.k_proj.weight")])
def test_parses_k_quant_snapshots(lib, tmp_path, dt, fill, name):
    """parses_q4k_snapshot / parses_q6k_snapshot: like DsqReader::open, the payload length of a block dtype is not
    compared with the dims at open time (4096 bytes are not 32 x 2 blocks); the engine checks it when uploading."""
    q = bytes([fill]) * 4096
    st, msg, hdr, recs = inspect(lib, tmp_path, build_snapshot_bytes(dt, dt, 256, name, 32, 512, q))
    assert st == 0, msg
    assert recs[0].q_dtype == dt and recs[0].in_dim == 512 and recs[0].q_len == 4096 and recs[0].first_q_byte == fill
    assert hdr.block_size == 256


def test_parses_f32_snapshot(lib, tmp_path):
    q = bytes([0x11]) * (2 * 3 * 4)
    st, msg, _, recs = inspect(lib, tmp_path, build_snapshot_bytes(Q8_0, F32, 32, "float.weight", 2, 3, q))
    assert st == 0, msg
    assert recs[0].q_dtype == F32 and recs[0].in_dim == 3 and recs[0].q_len == len(q)


def test_rejects_float_with_wrong_byte_len(lib, tmp_path):
    q = bytes([0x22]) * (2 * 3 * 4 - 1)
    st, msg, _, _ = inspect(lib, tmp_path, build_snapshot_bytes(Q8_0, F32, 32, "bad.float", 2, 3, q))
    assert st != 0 and "snapshot validation failed" in msg and "expected" in msg


def test_header_errors_use_the_reference_wording(lib, tmp_path):
    """DsqError's Display strings (lib.rs:18-33) for the header-level failures."""
    q = bytes(64 // 32 * 34 * 4)
    ok = dict(header_dtype=Q8_0, record_dtype=Q8_0, block_size=32, name="w", out_dim=4, in_dim=64, q_bytes=q)
    st, msg, _, _ = inspect(lib, tmp_path, build_snapshot_bytes(**ok, magic=b"DSQSNAX"))
    assert st != 0 and "invalid snapshot magic" in msg
    st, msg, _, _ = inspect(lib, tmp_path, build_snapshot_bytes(**ok, version=2))
    assert st != 0 and "unsupported snapshot version 2, expected 1" in msg
    st, msg, _, _ = inspect(lib, tmp_path, build_snapshot_bytes(**{**ok, "block_size": 0}))
    assert st != 0 and "block_size must be non-zero" in msg
    st, msg, _, _ = inspect(lib, tmp_path, build_snapshot_bytes(**{**ok, "block_size": 256}))
    assert st != 0 and "mismatches expected" in msg
    st, msg, _, _ = inspect(lib, tmp_path, build_snapshot_bytes(**{**ok, "record_dtype": 7}))
    assert st != 0 and "snapshot malformed" in msg and "unsupported tensor dtype code 7" in msg
    st, msg, _, _ = inspect(lib, tmp_path, build_snapshot_bytes(**ok)[:-5])
    assert st != 0 and "exceeds file size" in msg
    st, msg, _, _ = inspect(lib, tmp_path, build_snapshot_bytes(**{**ok, "q_bytes": b""}))
    assert st != 0 and "empty quantized payload" in msg


def test_real_snapshot_written_by_the_oracle_is_accepted(lib, tmp_path):
    """A full model snapshot (exporter dtype assignment) of the tiny config: every record is listed with its dims."""
    from oracle import dsq
    from tests.helpers import tiny_model

    cfg, ck, d = tiny_model("bf16")
    snap = str(tmp_path / "model.q4k.dsq")
    assigned = dsq.write_model_snapshot(snap, cfg, ck, dsq.Q4K)
    hdr = Hdr()
    assert lib.dsocr_dsq_inspect(snap.encode(), C.byref(hdr), None, 0) == 0, lib.dsocr_last_error().decode()
    assert hdr.tensor_count == len(assigned) and hdr.default_qdtype == Q4K and hdr.block_size == 256
    recs = (Rec * hdr.tensor_count)()
    assert lib.dsocr_dsq_inspect(snap.encode(), C.byref(hdr), recs, hdr.tensor_count) == 0
    seen = {r.name.decode(): r for r in recs}
    assert set(seen) == set(assigned)
    for name, dt in assigned.items():
        r = seen[name]
        assert r.q_dtype == dt
        assert r.q_len == r.out_dim * (r.in_dim // dsq.BLOCK[dt]) * dsq.BLOCK_BYTES[dt]
