"""DSQ snapshot container + ggml block formats (Q8_0 / Q4_K / Q6_K).  Test infrastructure only.

Follows:
  * crates/dsq/src/lib.rs:14-15, 60-110, 314-391 (container: magic DSQSNAP, v1, LE header + records + payload)
  * crates/dsq-writer/src/lib.rs:555-598 (Q8_0 quantiser: d = amax/127 in f32 stored as f16, q = round-half-away(v / d_f32))
  * crates/dsq-models/src/adapters/deepseek_ocr.rs:41-154 + crates/dsq-cli/src/main.rs:953-998 (which tensor gets which
    dtype; fallback chain K-quant -> Q8_0 -> float when in_dim % block != 0; lm_head / projector forced to Q8_0)
Block layouts are ggml's (the reference gets them from candle's k_quants).  Dequantisation is pinned against
`gguf.quants.dequantize` in tests/test_oracle_pins.py.  The K-quant *quantisers* below are NOT ggml's search-based ones
(candle `BlockQ4K::from_float` is not available offline): they are simple min/max quantisers that emit valid blocks,
which is all the dequant-matmul path needs (parity contract: y = x . dequant(W)^T, SURVEY.md 8c "Parity note for DSQ").
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

MAGIC = b"DSQSNAP"
VERSION = 1
Q8_0, Q4K, Q6K, F16, BF16, F32 = 8, 12, 14, 1, 16, 0
BLOCK = {Q8_0: 32, Q4K: 256, Q6K: 256}
BLOCK_BYTES = {Q8_0: 34, Q4K: 144, Q6K: 210}
BIAS_F32 = 4


# ---------------------------------------------------------------------------------------------- quantisers
def quantize_q8_0(w: np.ndarray) -> bytes:
    """dsq-writer/src/lib.rs:555-598.  w: [rows, cols] f32, cols % 32 == 0."""
    rows, cols = w.shape
    assert cols % 32 == 0
    blk = w.astype(np.float32).reshape(rows, cols // 32, 32)
    amax = np.abs(blk).max(-1)
    scale = np.where(amax > 0, amax / np.float32(127.0), np.float32(0.0)).astype(np.float32)
    inv = np.where(scale > 0, np.float32(1.0) / np.where(scale > 0, scale, 1), 0).astype(np.float32)
    v = blk * inv[..., None]
    q = np.clip(np.sign(v) * np.floor(np.abs(v) + np.float32(0.5)), -128, 127).astype(np.int8)  # f32::round
    out = np.zeros((rows, cols // 32, 34), dtype=np.uint8)
    out[..., :2] = scale.astype(np.float16).view(np.uint8).reshape(rows, cols // 32, 2)
    out[..., 2:] = q.view(np.uint8)
    return out.tobytes()


def _pack_q4k_scales(sc: np.ndarray, mn: np.ndarray) -> np.ndarray:
    """6-bit scales / mins for 8 sub-blocks -> 12 bytes (ggml get_scale_min_k4 layout)."""
    out = np.zeros(sc.shape[:-1] + (12,), dtype=np.uint8)
    for j in range(4):
        out[..., j] = (sc[..., j] & 63) | ((sc[..., j + 4] >> 4) << 6)
        out[..., j + 4] = (mn[..., j] & 63) | ((mn[..., j + 4] >> 4) << 6)
        out[..., j + 8] = (sc[..., j + 4] & 0xF) | ((mn[..., j + 4] & 0xF) << 4)
    return out


def quantize_q4k(w: np.ndarray) -> bytes:
    """Valid Q4_K blocks {f16 d; f16 dmin; u8 scales[12]; u8 qs[128]} (simple min/max quantiser)."""
    rows, cols = w.shape
    assert cols % 256 == 0
    x = w.astype(np.float32).reshape(rows, cols // 256, 8, 32)
    mins = np.minimum(x.min(-1), 0.0)
    maxs = x.max(-1)
    sub_scale = (maxs - mins) / 15.0            # per sub-block step
    sub_min = -mins                             # value = d*sc*q - dmin*m
    d = sub_scale.max(-1) / 63.0
    dmin = sub_min.max(-1) / 63.0
    d16 = d.astype(np.float16)
    dmin16 = dmin.astype(np.float16)
    df, dmf = d16.astype(np.float32), dmin16.astype(np.float32)
    sc = np.where(df[..., None] > 0, np.round(sub_scale / np.where(df[..., None] > 0, df[..., None], 1)), 0).clip(0, 63).astype(np.uint8)
    mn = np.where(dmf[..., None] > 0, np.round(sub_min / np.where(dmf[..., None] > 0, dmf[..., None], 1)), 0).clip(0, 63).astype(np.uint8)
    eff_s = df[..., None] * sc
    eff_m = dmf[..., None] * mn
    q = np.where(eff_s[..., None] > 0, np.round((x + eff_m[..., None]) / np.where(eff_s[..., None] > 0, eff_s[..., None], 1)), 0)
    q = q.clip(0, 15).astype(np.uint8)          # [rows, nb, 8, 32]
    qs = np.zeros((rows, cols // 256, 4, 32), dtype=np.uint8)
    for g in range(4):                          # byte l of group g: low nibble = sub-block 2g, high = 2g+1
        qs[..., g, :] = q[..., 2 * g, :] | (q[..., 2 * g + 1, :] << 4)
    out = np.zeros((rows, cols // 256, 144), dtype=np.uint8)
    out[..., 0:2] = d16.view(np.uint8).reshape(rows, -1, 2)
    out[..., 2:4] = dmin16.view(np.uint8).reshape(rows, -1, 2)
    out[..., 4:16] = _pack_q4k_scales(sc, mn)
    out[..., 16:144] = qs.reshape(rows, cols // 256, 128)
    return out.tobytes()


def quantize_q6k(w: np.ndarray) -> bytes:
    """Valid Q6_K blocks {u8 ql[128]; u8 qh[64]; i8 scales[16]; f16 d} (simple symmetric quantiser)."""
    rows, cols = w.shape
    assert cols % 256 == 0
    x = w.astype(np.float32).reshape(rows, cols // 256, 16, 16)
    amax = np.abs(x).max(-1)                     # per 16-weight group
    gscale = amax / 31.0
    d = gscale.max(-1) / 127.0
    d16 = d.astype(np.float16)
    df = d16.astype(np.float32)
    sc = np.where(df[..., None] > 0, np.round(gscale / np.where(df[..., None] > 0, df[..., None], 1)), 0).clip(-128, 127).astype(np.int8)
    eff = df[..., None] * sc.astype(np.float32)
    q = np.where(eff[..., None] != 0, np.round(x / np.where(eff[..., None] != 0, eff[..., None], 1)), 0).clip(-32, 31).astype(np.int32) + 32
    q = q.reshape(rows, cols // 256, 256).astype(np.uint8)   # 0..63
    ql = np.zeros((rows, cols // 256, 128), dtype=np.uint8)
    qh = np.zeros((rows, cols // 256, 64), dtype=np.uint8)
    for half in range(2):                        # ggml dequantize_row_q6_K index pattern
        base = half * 128
        for l in range(32):
            q1, q2, q3, q4 = (q[..., base + l], q[..., base + 32 + l], q[..., base + 64 + l], q[..., base + 96 + l])
            ql[..., half * 64 + l] = (q1 & 0xF) | ((q3 & 0xF) << 4)
            ql[..., half * 64 + 32 + l] = (q2 & 0xF) | ((q4 & 0xF) << 4)
            qh[..., half * 32 + l] = (q1 >> 4) | ((q2 >> 4) << 2) | ((q3 >> 4) << 4) | ((q4 >> 4) << 6)
    out = np.zeros((rows, cols // 256, 210), dtype=np.uint8)
    out[..., 0:128] = ql
    out[..., 128:192] = qh
    out[..., 192:208] = sc.view(np.uint8)
    out[..., 208:210] = d16.view(np.uint8).reshape(rows, -1, 2)
    return out.tobytes()


# ---------------------------------------------------------------------------------------------- dequantisers
def dequantize(data: bytes, dtype: int, rows: int, cols: int) -> np.ndarray:
    """-> f32 [rows, cols].  ggml dequantize_row_{q8_0,q4_K,q6_K} semantics."""
    raw = np.frombuffer(data, dtype=np.uint8)
    if dtype == F32:
        return raw.view(np.float32).reshape(rows, cols).copy()
    if dtype == F16:
        return raw.view(np.float16).astype(np.float32).reshape(rows, cols)
    if dtype == BF16:
        u = raw.view(np.uint16).astype(np.uint32) << 16
        return u.view(np.float32).reshape(rows, cols)
    nb = cols // BLOCK[dtype]
    blk = raw.reshape(rows, nb, BLOCK_BYTES[dtype])
    if dtype == Q8_0:
        d = blk[..., :2].copy().view(np.float16).astype(np.float32)  # [rows, nb, 1]
        q = blk[..., 2:].view(np.int8).astype(np.float32)
        return (d * q).reshape(rows, cols)
    if dtype == Q4K:
        d = blk[..., 0:2].copy().view(np.float16).astype(np.float32)[..., 0]
        dmin = blk[..., 2:4].copy().view(np.float16).astype(np.float32)[..., 0]
        s = blk[..., 4:16]
        sc = np.zeros((rows, nb, 8), dtype=np.float32)
        mn = np.zeros((rows, nb, 8), dtype=np.float32)
        for j in range(4):
            sc[..., j] = s[..., j] & 63
            mn[..., j] = s[..., j + 4] & 63
            sc[..., j + 4] = (s[..., j + 8] & 0xF) | ((s[..., j] >> 6) << 4)
            mn[..., j + 4] = (s[..., j + 8] >> 4) | ((s[..., j + 4] >> 6) << 4)
        qs = blk[..., 16:144].reshape(rows, nb, 4, 32)
        out = np.zeros((rows, nb, 8, 32), dtype=np.float32)
        for g in range(4):
            out[..., 2 * g, :] = d[..., None] * sc[..., 2 * g, None] * (qs[..., g, :] & 0xF) - dmin[..., None] * mn[..., 2 * g, None]
            out[..., 2 * g + 1, :] = d[..., None] * sc[..., 2 * g + 1, None] * (qs[..., g, :] >> 4) - dmin[..., None] * mn[..., 2 * g + 1, None]
        return out.reshape(rows, cols)
    if dtype == Q6K:
        ql = blk[..., 0:128].astype(np.int32)
        qh = blk[..., 128:192].astype(np.int32)
        sc = blk[..., 192:208].view(np.int8).astype(np.float32)
        d = blk[..., 208:210].copy().view(np.float16).astype(np.float32)[..., 0]
        out = np.zeros((rows, nb, 256), dtype=np.float32)
        for half in range(2):
            base = half * 128
            for l in range(32):
                isx = l // 16
                q1 = ((ql[..., half * 64 + l] & 0xF) | (((qh[..., half * 32 + l] >> 0) & 3) << 4)) - 32
                q2 = ((ql[..., half * 64 + 32 + l] & 0xF) | (((qh[..., half * 32 + l] >> 2) & 3) << 4)) - 32
                q3 = ((ql[..., half * 64 + l] >> 4) | (((qh[..., half * 32 + l] >> 4) & 3) << 4)) - 32
                q4 = ((ql[..., half * 64 + 32 + l] >> 4) | (((qh[..., half * 32 + l] >> 6) & 3) << 4)) - 32
                out[..., base + l] = d * sc[..., half * 8 + isx + 0] * q1
                out[..., base + 32 + l] = d * sc[..., half * 8 + isx + 2] * q2
                out[..., base + 64 + l] = d * sc[..., half * 8 + isx + 4] * q3
                out[..., base + 96 + l] = d * sc[..., half * 8 + isx + 6] * q4
        return out.reshape(rows, cols)
    raise ValueError(f"unsupported dtype {dtype}")


# ---------------------------------------------------------------------------------------------- container
@dataclass
class Record:
    name: str
    out_dim: int
    in_dim: int
    q_dtype: int
    q_offset: int
    q_len: int
    bias_offset: Optional[int] = None
    bias_len: Optional[int] = None
    bias_dtype: Optional[int] = None


def _wstr(s: str) -> bytes:
    b = s.encode()
    return struct.pack("<I", len(b)) + b


def write_snapshot(path: str, default_dtype: int, tensors: List[Tuple[str, int, int, int, bytes, Optional[bytes]]],
                   model_id: str = "deepseek-ocr", backend: str = "CPU", candle_version: str = "0.9.2") -> None:
    """tensors: (name, out_dim, in_dim, q_dtype, q_bytes, bias_f32_bytes|None).  Layout of dsq/src/lib.rs:314-391."""
    head = MAGIC + struct.pack("<I", VERSION) + _wstr(candle_version) + _wstr(model_id) + _wstr(backend)
    head += struct.pack("<III", default_dtype, BLOCK[default_dtype], len(tensors))
    meta_len = len(head) + sum(4 + len(n.encode()) + 12 + 32 + 4 for n, *_ in tensors)
    off = meta_len
    recs = b""
    payload = []
    for name, out_dim, in_dim, dt, q, bias in tensors:
        q_off = off
        off += len(q)
        if bias is not None:
            b_off, b_len, b_dt = off, len(bias), BIAS_F32
            off += len(bias)
        else:
            b_off = b_len = b_dt = 0
        recs += _wstr(name) + struct.pack("<III", out_dim, in_dim, dt) + struct.pack("<QQQQ", q_off, len(q), b_off, b_len)
        recs += struct.pack("<I", b_dt)
        payload.append(q)
        if bias is not None:
            payload.append(bias)
    with open(path, "wb") as f:
        f.write(head + recs)
        for p in payload:
            f.write(p)


def read_snapshot(path: str):
    """-> (header dict, {name: Record}, bytes).  parse_index + validations of dsq/src/lib.rs."""
    data = open(path, "rb").read()
    if data[:7] != MAGIC:
        raise ValueError(f"invalid snapshot magic: found {data[:7]!r}")
    pos = 7
    (ver,) = struct.unpack_from("<I", data, pos); pos += 4
    if ver != VERSION:
        raise ValueError(f"unsupported snapshot version {ver}, expected {VERSION}")

    def rstr():
        nonlocal pos
        (n,) = struct.unpack_from("<I", data, pos); pos += 4
        s = data[pos:pos + n].decode(); pos += n
        return s

    hdr = {"candle_version": rstr(), "model_id": rstr(), "backend": rstr()}
    hdr["default_qdtype"], hdr["block_size"], count = struct.unpack_from("<III", data, pos); pos += 12
    if hdr["block_size"] != BLOCK.get(hdr["default_qdtype"]):
        raise ValueError("snapshot block size mismatches dtype")
    recs: Dict[str, Record] = {}
    for _ in range(count):
        name = rstr()
        out_dim, in_dim, dt = struct.unpack_from("<III", data, pos); pos += 12
        q_off, q_len, b_off, b_len = struct.unpack_from("<QQQQ", data, pos); pos += 32
        (b_dt,) = struct.unpack_from("<I", data, pos); pos += 4
        r = Record(name, out_dim, in_dim, dt, q_off, q_len)
        if b_len:
            r.bias_offset, r.bias_len, r.bias_dtype = b_off, b_len, b_dt
        if name in recs:
            raise ValueError(f"duplicate tensor record `{name}`")
        recs[name] = r
    for r in recs.values():
        if r.q_len == 0 or r.q_offset < pos or r.q_offset + r.q_len > len(data):
            raise ValueError(f"tensor `{r.name}` payload out of bounds")
        if r.q_dtype in BLOCK and r.in_dim % BLOCK[r.q_dtype]:
            raise ValueError(f"tensor `{r.name}` in_dim {r.in_dim} not divisible by block_size")
    return hdr, recs, data


# ---------------------------------------------------------------------------------------------- model snapshot
def linear_specs(cfg) -> List[Tuple[str, int, int]]:
    """dsq-models/src/adapters/deepseek_ocr.rs:41-139 (text scope: decoder linears + lm_head)."""
    H = cfg.hidden_size
    specs = []
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        for n in "qkvo":
            specs.append((p + f"self_attn.{n}_proj.weight", H, H))

        def mlp(prefix, inter):
            return [(prefix + "gate_proj.weight", inter, H), (prefix + "up_proj.weight", inter, H),
                    (prefix + "down_proj.weight", H, inter)]
        if i < cfg.first_k_dense_replace:
            specs += mlp(p + "mlp.", cfg.intermediate_size)
        else:
            for e in range(cfg.n_routed_experts):
                specs += mlp(f"{p}mlp.experts.{e}.", cfg.moe_intermediate_size)
            specs += mlp(p + "mlp.shared_experts.", cfg.moe_intermediate_size * cfg.n_shared_experts)
    specs.append(("lm_head.weight", cfg.vocab_size, H))
    return specs


def choose_dtype(name: str, in_dim: int, primary: int) -> int:
    """adapter recommend_dtype (:141-154) + dsq-cli fallback chain (main.rs:953-998)."""
    want = primary
    if primary != Q8_0 and name in ("lm_head.weight", "model.projector.layers.weight"):
        want = Q8_0
    if in_dim % BLOCK[want] == 0:
        return want
    if in_dim % 32 == 0:
        return Q8_0
    return BF16


_QUANT = {Q8_0: quantize_q8_0, Q4K: quantize_q4k, Q6K: quantize_q6k}


def write_model_snapshot(path: str, cfg, ckpt, primary: int) -> Dict[str, int]:
    """Synthesise a q8_0 / q4k / q6k snapshot of the decoder from a checkpoint; returns {tensor: dtype}."""
    import torch

    tensors, assigned = [], {}
    for name, out_dim, in_dim in linear_specs(cfg):
        w = ckpt[name].to(torch.float32).numpy()
        dt = choose_dtype(name, in_dim, primary)
        if dt in _QUANT:
            q = _QUANT[dt](w)
        else:
            q = ckpt[name].to(torch.bfloat16).view(torch.int16).numpy().tobytes()
        tensors.append((name, out_dim, in_dim, dt, q, None))
        assigned[name] = dt
    write_snapshot(path, primary, tensors)
    return assigned


def dequantized_checkpoint(path: str, ckpt) -> dict:
    """Checkpoint with every snapshot tensor replaced by its f32 dequantisation (the DSQ parity oracle weights)."""
    import torch

    _, recs, data = read_snapshot(path)
    out = dict(ckpt)
    for name, r in recs.items():
        w = dequantize(data[r.q_offset:r.q_offset + r.q_len], r.q_dtype, r.out_dim, r.in_dim)
        out[name] = torch.from_numpy(np.ascontiguousarray(w))
    return out
