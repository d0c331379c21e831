"""DSQ snapshot exporter: what `deepseek-ocr-cli weights snapshot --in <safetensors> --out <path> --dtype q8_0|q4k|q6k
--targets text|text+projector` does (crates/dsq-cli/src/main.rs), on top of the library's writer (dsocr_dsq_writer_*).

  * which tensors: DeepSeek-OCR adapter, dsq-models/src/adapters/deepseek_ocr.rs:41-139 (per layer q/k/v/o, the dense MLP
    or every routed expert + the fused shared experts, lm_head, optionally the projector; optional `.bias` companions)
  * which dtype: adapter override (:141-154: lm_head / projector -> Q8_0 unless the primary is Q8_0), then the fallback
    chain of dsq-cli (main.rs:953-1004: Q4_K / Q6_K -> Q8_0 when in_dim is not a multiple of the block), then a float
    payload in the tensor's own dtype (main.rs:637-660)
The Q4_K / Q6_K quantisers of the library are ggml-style and not byte-pinned to candle's (see csrc/dsq_writer.cpp)."""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, List, Optional, Tuple

import numpy as np

from .binding import check, lib

Q8_0, Q4K, Q6K, F16, BF16, F32 = 8, 12, 14, 1, 16, 0
BLOCK = {Q8_0: 32, Q4K: 256, Q6K: 256}
DTYPE_NAMES = {"q8_0": Q8_0, "q4k": Q4K, "q4_k": Q4K, "q6k": Q6K, "q6_k": Q6K}
PROJECTOR = "model.projector.layers.weight"


def _lang(cfg: dict) -> dict:
    """language_config merged over the top level (adapter `language_config`)."""
    lc = cfg.get("language_config")
    return lc if isinstance(lc, dict) else cfg


def linear_specs(cfg: dict, include_projector: bool = False) -> List[Tuple[str, int, int, Optional[str]]]:
    """(weight name, out_dim, in_dim, bias name or None), in the adapter's order."""
    lc = _lang(cfg)
    H = int(lc["hidden_size"])
    layers = int(lc["num_hidden_layers"])
    heads = int(lc["num_attention_heads"])
    kv_heads = int(lc.get("num_key_value_heads") or heads)
    head_dim = H // heads
    v_head_dim = int(lc.get("v_head_dim") or 0) or head_dim
    inter = int(lc["intermediate_size"])
    moe_inter = int(lc.get("moe_intermediate_size") or 0) or inter
    n_routed = int(lc.get("n_routed_experts") or 0)
    n_shared = int(lc.get("n_shared_experts") or 0)
    moe_freq = int(lc.get("moe_layer_freq") or 0) or 1
    first_dense = int(lc.get("first_k_dense_replace") or 0)
    vocab = int(lc["vocab_size"])

    def mlp(prefix: str, i: int):
        return [(f"{prefix}.gate_proj.weight", i, H, None), (f"{prefix}.up_proj.weight", i, H, None),
                (f"{prefix}.down_proj.weight", H, i, None)]

    specs: List[Tuple[str, int, int, Optional[str]]] = []
    for l in range(layers):
        a = f"model.layers.{l}.self_attn"
        specs += [(f"{a}.q_proj.weight", heads * head_dim, H, f"{a}.q_proj.bias"),
                  (f"{a}.k_proj.weight", kv_heads * head_dim, H, f"{a}.k_proj.bias"),
                  (f"{a}.v_proj.weight", kv_heads * v_head_dim, H, f"{a}.v_proj.bias"),
                  (f"{a}.o_proj.weight", H, heads * v_head_dim, f"{a}.o_proj.bias")]
        m = f"model.layers.{l}.mlp"
        use_moe = n_routed > 0 and l >= first_dense and l % moe_freq == 0  # should_use_moe (weights.rs:609-619)
        if use_moe:
            for e in range(n_routed):
                specs += mlp(f"{m}.experts.{e}", moe_inter)
            if n_shared > 0:
                specs += mlp(f"{m}.shared_experts", moe_inter * n_shared)
        else:
            specs += mlp(m, inter)
    if lc.get("lm_head", True):
        specs.append(("lm_head.weight", vocab, H, None))
    if include_projector:
        pc = cfg["projector_config"]
        specs.append((PROJECTOR, int(pc["n_embed"]), int(pc["input_dim"]), "model.projector.layers.bias"))
    return specs


def select_dtype(name: str, in_dim: int, primary: int) -> Optional[int]:
    """-> block dtype, or None when the tensor must fall back to a float payload."""
    want = primary
    if primary != Q8_0 and name in ("lm_head.weight", PROJECTOR):
        want = Q8_0
    while True:
        if in_dim % BLOCK[want] == 0:
            return want
        if want in (Q4K, Q6K):
            want = Q8_0
            continue
        return None


def _float_code(np_dtype) -> int:
    return {"float32": F32, "float16": F16}.get(str(np_dtype), BF16)


def export_snapshot(config_path: str, safetensors_path: str, out_path: str, dtype: str = "q8_0", targets: str = "text",
                    model_id: str = "deepseek-ocr", backend: str = "CPU", candle_version: str = "dsocr-b200") -> Dict[str, int]:
    """Writes `<out_path with .dsq>`; returns {tensor name: dtype code written}."""
    import torch
    from safetensors import safe_open

    if dtype.lower() not in DTYPE_NAMES:
        raise ValueError(f"unsupported snapshot dtype `{dtype}` (q8_0, q4k, q6k)")
    if targets not in ("text", "text+projector"):
        raise ValueError(f"unsupported targets `{targets}`")
    primary = DTYPE_NAMES[dtype.lower()]
    cfg = json.load(open(config_path))
    L = lib()
    fp, u8 = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    L.dsocr_dsq_writer_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
    L.dsocr_dsq_writer_add_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, fp, fp]
    L.dsocr_dsq_writer_finalize.argtypes = [C.c_void_p]
    L.dsocr_dsq_writer_destroy.argtypes = [C.c_void_p]
    h = C.c_void_p()
    check(L.dsocr_dsq_writer_create(out_path.encode(), candle_version.encode(), model_id.encode(), backend.encode(), primary,
                                    C.byref(h)), "snapshot writer")
    written: Dict[str, int] = {}
    try:
        with safe_open(safetensors_path, framework="pt") as st:
            names = set(st.keys())
            for name, out_dim, in_dim, bias_name in linear_specs(cfg, targets == "text+projector"):
                if name not in names:
                    raise KeyError(f"checkpoint is missing tensor `{name}`")
                t = st.get_tensor(name)
                if tuple(t.shape) != (out_dim, in_dim):
                    raise ValueError(f"tensor `{name}` has shape {tuple(t.shape)}, expected ({out_dim}, {in_dim})")
                sel = select_dtype(name, in_dim, primary)
                code = sel if sel is not None else _float_code(str(t.dtype).replace("torch.", ""))
                w = np.ascontiguousarray(t.to(torch.float32).numpy())
                bias = None
                if bias_name and bias_name in names:
                    bias = np.ascontiguousarray(st.get_tensor(bias_name).to(torch.float32).numpy())
                check(L.dsocr_dsq_writer_add_tensor(h, name.encode(), out_dim, in_dim, code, w.ctypes.data_as(fp),
                                                    bias.ctypes.data_as(fp) if bias is not None else None), "snapshot writer")
                written[name] = code
        check(L.dsocr_dsq_writer_finalize(h), "snapshot writer")  # frees the handle
        h = None
    finally:
        if h is not None:
            L.dsocr_dsq_writer_destroy(h)
    return written
