"""Python mirror of the reference's engine interface (`load_model` / `OcrEngine`,
crates/core/src/inference.rs:162-209, crates/infer-deepseek/src/model/mod.rs:90-115) over the C ABI.

Only marshals host buffers; all compute happens in libdsocr.so on the GPU.  Error strings keep the
reference's context prefixes ("vision input failed", "image embedding failed", "prompt formatting failed").
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from .binding import DsocrError, check, lib

F32, F16, BF16 = 0, 1, 2
_DTYPES = {"f32": F32, "f16": F16, "bf16": BF16, "float16": F16, "bfloat16": BF16, "float32": F32}


class VisionSettingsC(C.Structure):
    _fields_ = [("base_size", C.c_uint32), ("image_size", C.c_uint32), ("crop_mode", C.c_int32)]


class DecodeParamsC(C.Structure):
    _fields_ = [("max_new_tokens", C.c_uint32), ("do_sample", C.c_int32), ("repetition_penalty", C.c_float),
                ("no_repeat_ngram_size", C.c_uint32), ("eos_token_id", C.c_int64), ("use_cache", C.c_int32),
                ("has_seed", C.c_int32), ("temperature", C.c_double), ("top_p", C.c_double), ("top_k", C.c_uint32),
                ("reserved_", C.c_uint32), ("seed", C.c_uint64)]


class RequestC(C.Structure):
    _fields_ = [("n_images", C.c_int32), ("rgb", C.POINTER(C.POINTER(C.c_uint8))), ("widths", C.POINTER(C.c_int32)),
                ("heights", C.POINTER(C.c_int32)), ("n_segments", C.c_int32),
                ("segments", C.POINTER(C.POINTER(C.c_int64))), ("segment_lens", C.POINTER(C.c_int32))]


class EngineInfoC(C.Structure):
    _fields_ = [("device_ordinal", C.c_int32), ("dtype", C.c_int32), ("sm_count", C.c_int32),
                ("hidden_size", C.c_int32), ("num_layers", C.c_int32), ("vocab_size", C.c_int32),
                ("n_routed_experts", C.c_int32), ("quantized", C.c_int32), ("device_name", C.c_char * 64)]


TOKEN_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_size_t, C.POINTER(C.c_int64))


@dataclass
class VisionSettings:  # crates/core/src/inference.rs:10-16
    base_size: int = 1024
    image_size: int = 640
    crop_mode: bool = True

    def c(self) -> VisionSettingsC:
        return VisionSettingsC(self.base_size, self.image_size, 1 if self.crop_mode else 0)


@dataclass
class DecodeParameters:  # crates/core/src/inference.rs:18-34, defaults :66-78
    max_new_tokens: int = 512
    do_sample: bool = False
    repetition_penalty: float = 1.0
    no_repeat_ngram_size: Optional[int] = 20
    eos_token_id: Optional[int] = 1
    use_cache: bool = True
    temperature: float = 0.0
    top_p: Optional[float] = 1.0
    top_k: Optional[int] = None
    seed: Optional[int] = None

    def c(self) -> DecodeParamsC:
        return DecodeParamsC(self.max_new_tokens, 1 if self.do_sample else 0, self.repetition_penalty,
                             self.no_repeat_ngram_size or 0, -1 if self.eos_token_id is None else self.eos_token_id,
                             1 if self.use_cache else 0, 0 if self.seed is None else 1, float(self.temperature),
                             -1.0 if self.top_p is None else float(self.top_p), int(self.top_k or 0), 0,
                             int(self.seed or 0))


@dataclass
class DecodeOutcome:  # crates/core/src/inference.rs:162-177 (text decoding stays with the host tokenizer)
    prompt_tokens: int
    response_tokens: int
    generated_tokens: List[int] = field(default_factory=list)


def _ptr_array(arrs, ctype):
    n = len(arrs)
    out = (C.POINTER(ctype) * n)()
    for i, a in enumerate(arrs):
        out[i] = a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None
    return out


class OcrEngine:
    """`Box<dyn OcrEngine>` for the DeepSeek-OCR model kind."""

    def __init__(self, config_path: str, weights_path: str, snapshot_path: Optional[str] = None, device: int = 0,
                 dtype: str = "bf16"):
        self._h = C.c_void_p()
        self._lib = lib()
        st = self._lib.dsocr_engine_create(config_path.encode(), weights_path.encode(),
                                           snapshot_path.encode() if snapshot_path else None, device,
                                           _DTYPES[dtype], C.byref(self._h))
        check(st, "load_model")
        self._lib.dsocr_launch_count.restype = C.c_longlong
        info = EngineInfoC()
        check(self._lib.dsocr_engine_info_get(self._h, C.byref(info)), "engine_info")
        self.info = info
        self.hidden = info.hidden_size
        self.vocab = info.vocab_size

    def close(self):
        if self._h:
            self._lib.dsocr_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- options / introspection ------------------------------------------------------------------
    def set_option(self, name: str, value: int):
        check(self._lib.dsocr_engine_set_option(self._h, name.encode(), int(value)), "set_option")

    def tap(self, name: str) -> np.ndarray:
        n = C.c_size_t()
        check(self._lib.dsocr_vision_tap(self._h, name.encode(), None, 0, C.byref(n)), "tap")
        out = np.empty(n.value, dtype=np.float32)
        check(self._lib.dsocr_vision_tap(self._h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), n.value,
                                         C.byref(n)), "tap")
        return out

    def moe_stats(self) -> tuple:
        """(non-empty (layer, expert) segments summed over decode steps, decode steps) since the last read."""
        v = (C.c_double * 2)()
        check(self._lib.dsocr_moe_stats(self._h, v), "moe_stats")
        return float(v[0]), float(v[1])

    def launch_count(self) -> int:
        return int(self._lib.dsocr_launch_count(self._h))

    def timings(self) -> dict:
        ms = (C.c_double * 5)()
        check(self._lib.dsocr_last_timings(self._h, ms, 5), "timings")
        names = ["vision.prepare_inputs", "vision.compute_embeddings", "decode.prefill", "decode.iterative",
                 "decode.generate"]
        return dict(zip(names, list(ms)))

    # -- preprocessing (host integer path of the library) -----------------------------------------
    def preprocess(self, rgb: np.ndarray, vs: VisionSettings):
        h, w = rgb.shape[:2]
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        g = vs.base_size if vs.crop_mode else vs.image_size
        gout = np.empty((g, g, 3), dtype=np.uint8)
        tiles = np.empty((9, vs.image_size, vs.image_size, 3), dtype=np.uint8)
        n, cw, ch = C.c_int(), C.c_int(), C.c_int()
        u8 = C.POINTER(C.c_uint8)
        check(self._lib.dsocr_preprocess(rgb.ctypes.data_as(u8), w, h, vs.c(), gout.ctypes.data_as(u8),
                                         tiles.ctypes.data_as(u8), C.byref(n), C.byref(cw), C.byref(ch)), "preprocess")
        return gout, tiles[: n.value].copy(), (cw.value, ch.value)

    def preprocess_gpu(self, rgb: np.ndarray, vs: VisionSettings):
        h, w = rgb.shape[:2]
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        g = vs.base_size if vs.crop_mode else vs.image_size
        gout = np.empty((g, g, 3), dtype=np.uint8)
        tiles = np.empty((9, vs.image_size, vs.image_size, 3), dtype=np.uint8)
        n, cw, ch = C.c_int(), C.c_int(), C.c_int()
        u8 = C.POINTER(C.c_uint8)
        check(self._lib.dsocr_preprocess_gpu(self._h, rgb.ctypes.data_as(u8), w, h, vs.c(), gout.ctypes.data_as(u8),
                                             tiles.ctypes.data_as(u8), C.byref(n), C.byref(cw), C.byref(ch)), "preprocess_gpu")
        return gout, tiles[: n.value].copy(), (cw.value, ch.value)

    # -- compute_image_embeddings -----------------------------------------------------------------
    def vision_encode(self, global_chw: np.ndarray, patches: Optional[np.ndarray], crop_shape) -> np.ndarray:
        g = np.ascontiguousarray(global_chw, dtype=np.float32)
        gs = g.shape[-1]
        n_p = 0 if patches is None else patches.shape[0]
        ps = 0 if patches is None else patches.shape[-1]
        p = None if patches is None else np.ascontiguousarray(patches, dtype=np.float32)
        cap = 4096
        out = np.empty((cap, self.hidden), dtype=np.float32)
        n_rows = C.c_int(cap)
        fp = C.POINTER(C.c_float)
        cw, ch = crop_shape if crop_shape else (1, 1)
        check(self._lib.dsocr_vision_encode(self._h, g.ctypes.data_as(fp), gs, p.ctypes.data_as(fp) if p is not None else None,
                                            n_p, ps, cw, ch, out.ctypes.data_as(fp), C.byref(n_rows)), "vision_encode")
        return out[: n_rows.value].copy()

    def vision_encode_u8(self, globals_u8: Sequence[np.ndarray], tiles_u8: Sequence[Optional[np.ndarray]],
                         crop_shapes: Sequence[tuple], patch_size: int) -> List[np.ndarray]:
        n = len(globals_u8)
        gs = globals_u8[0].shape[0]
        gl = [np.ascontiguousarray(g, dtype=np.uint8) for g in globals_u8]
        tl = [None if (t is None or len(t) == 0) else np.ascontiguousarray(t, dtype=np.uint8) for t in tiles_u8]
        nt = (C.c_int * n)(*[0 if t is None else t.shape[0] for t in tl])
        cw = (C.c_int * n)(*[c[0] for c in crop_shapes])
        ch = (C.c_int * n)(*[c[1] for c in crop_shapes])
        outs = [np.empty((4096, self.hidden), dtype=np.float32) for _ in range(n)]
        n_rows = (C.c_int * n)()
        check(self._lib.dsocr_vision_encode_u8_batch(self._h, n, _ptr_array(gl, C.c_uint8), gs, _ptr_array(tl, C.c_uint8),
                                                     nt, patch_size, cw, ch, _ptr_array(outs, C.c_float), n_rows),
              "vision_encode_u8_batch")
        return [o[: n_rows[i]].copy() for i, o in enumerate(outs)]

    # -- generate ---------------------------------------------------------------------------------
    def generate_batch(self, input_ids: Sequence[Sequence[int]], masks: Sequence[Sequence[int]],
                       image_rows: Sequence[Optional[np.ndarray]], params: DecodeParameters,
                       callback: Optional[Callable[[int, int, List[int]], None]] = None) -> List[List[int]]:
        n = len(input_ids)
        ids = [np.asarray(x, dtype=np.int64) for x in input_ids]
        ms = [np.asarray(x, dtype=np.uint8) for x in masks]
        rows = [None if r is None else np.ascontiguousarray(r, dtype=np.float32) for r in image_rows]
        nt = (C.c_int * n)(*[len(x) for x in ids])
        nr = (C.c_int * n)(*[0 if r is None else r.shape[0] for r in rows])
        outs = [np.zeros(max(1, params.max_new_tokens), dtype=np.int64) for _ in range(n)]
        n_out = (C.c_int * n)()
        cb = TOKEN_CB(lambda user, page, count, toks: callback(page, count, [toks[i] for i in range(count)])) if callback else None
        p = params.c()
        check(self._lib.dsocr_generate_batch(self._h, n, _ptr_array(ids, C.c_int64), _ptr_array(ms, C.c_uint8), nt,
                                             _ptr_array(rows, C.c_float), nr, C.byref(p),
                                             cb if cb else C.cast(None, TOKEN_CB), None, _ptr_array(outs, C.c_int64), n_out),
              "generate")
        return [outs[i][: n_out[i]].tolist() for i in range(n)]

    def generate_forced(self, input_ids, masks, image_rows, params: DecodeParameters, forced: Sequence[Sequence[int]],
                        want_logits: bool = False):
        n = len(input_ids)
        steps = len(forced[0])
        ids = [np.asarray(x, dtype=np.int64) for x in input_ids]
        ms = [np.asarray(x, dtype=np.uint8) for x in masks]
        rows = [None if r is None else np.ascontiguousarray(r, dtype=np.float32) for r in image_rows]
        fz = [np.asarray(x, dtype=np.int64) for x in forced]
        nt = (C.c_int * n)(*[len(x) for x in ids])
        nr = (C.c_int * n)(*[0 if r is None else r.shape[0] for r in rows])
        sel = [np.zeros(steps, dtype=np.int64) for _ in range(n)]
        lg = [np.zeros((steps, self.vocab), dtype=np.float32) if want_logits else None for _ in range(n)]
        p = params.c()
        check(self._lib.dsocr_generate_forced(self._h, n, _ptr_array(ids, C.c_int64), _ptr_array(ms, C.c_uint8), nt,
                                              _ptr_array(rows, C.c_float), nr, C.byref(p), _ptr_array(fz, C.c_int64),
                                              steps, _ptr_array(sel, C.c_int64), _ptr_array(lg, C.c_float)),
              "generate_forced")
        return [s.tolist() for s in sel], lg

    # -- OcrEngine::decode minus tokenizer ----------------------------------------------------------
    def decode_pages(self, pages_rgb: Sequence[np.ndarray], vs: VisionSettings, seg0: Sequence[int], seg1: Sequence[int],
                     image_token_id: int, params: DecodeParameters, callback=None) -> List[DecodeOutcome]:
        n = len(pages_rgb)
        imgs = [np.ascontiguousarray(p, dtype=np.uint8) for p in pages_rgb]
        ws = (C.c_int * n)(*[p.shape[1] for p in imgs])
        hs = (C.c_int * n)(*[p.shape[0] for p in imgs])
        s0 = np.asarray(seg0, dtype=np.int64)
        s1 = np.asarray(seg1, dtype=np.int64)
        outs = [np.zeros(max(1, params.max_new_tokens), dtype=np.int64) for _ in range(n)]
        n_out = (C.c_int * n)()
        n_prompt = (C.c_int * n)()
        cb = TOKEN_CB(lambda user, page, count, toks: callback(page, count, [toks[i] for i in range(count)])) if callback else None
        p = params.c()
        i64 = C.POINTER(C.c_int64)
        check(self._lib.dsocr_decode_pages(self._h, n, _ptr_array(imgs, C.c_uint8), ws, hs, vs.c(),
                                           s0.ctypes.data_as(i64), len(s0), s1.ctypes.data_as(i64), len(s1),
                                           C.c_int64(image_token_id), C.byref(p), cb if cb else C.cast(None, TOKEN_CB), None,
                                           _ptr_array(outs, C.c_int64), n_out, n_prompt), "decode")
        return [DecodeOutcome(n_prompt[i], n_out[i], outs[i][: n_out[i]].tolist()) for i in range(n)]


    def decode_requests(self, requests: Sequence[tuple], vs: VisionSettings, image_token_id: int,
                        params: DecodeParameters, callback=None) -> List[DecodeOutcome]:
        """`OcrEngine::decode` for a batch of requests; request = (images: list of RGB8 arrays, segments: list of
        token-id lists around the <image> slots, len(images) + 1 of them)."""
        n = len(requests)
        reqs = (RequestC * n)()
        keep = []
        for r, (images, segments) in enumerate(requests):
            imgs = [np.ascontiguousarray(p, dtype=np.uint8) for p in images]
            segs = [np.asarray(x, dtype=np.int64) for x in segments]
            ws = (C.c_int32 * max(1, len(imgs)))(*[p.shape[1] for p in imgs])
            hs = (C.c_int32 * max(1, len(imgs)))(*[p.shape[0] for p in imgs])
            lens = (C.c_int32 * max(1, len(segs)))(*[len(x) for x in segs])
            ip, sp = _ptr_array(imgs, C.c_uint8), _ptr_array(segs, C.c_int64)
            keep.append((imgs, segs, ws, hs, lens, ip, sp))
            reqs[r] = RequestC(len(imgs), C.cast(ip, C.POINTER(C.POINTER(C.c_uint8))), ws, hs, len(segs),
                               C.cast(sp, C.POINTER(C.POINTER(C.c_int64))), lens)
        outs = [np.zeros(max(1, params.max_new_tokens), dtype=np.int64) for _ in range(n)]
        n_out = (C.c_int * n)()
        n_prompt = (C.c_int * n)()
        cb = TOKEN_CB(lambda user, page, count, toks: callback(page, count, [toks[i] for i in range(count)])) if callback else None
        p = params.c()
        check(self._lib.dsocr_decode_requests(self._h, n, reqs, vs.c(), C.c_int64(image_token_id), C.byref(p),
                                              cb if cb else C.cast(None, TOKEN_CB), None, _ptr_array(outs, C.c_int64),
                                              n_out, n_prompt), "decode")
        return [DecodeOutcome(n_prompt[i], n_out[i], outs[i][: n_out[i]].tolist()) for i in range(n)]

    # -- staged variant (device-resident pages) ---------------------------------------------------
    def stage_pages(self, pages_rgb: Sequence[np.ndarray], vs: VisionSettings):
        n = len(pages_rgb)
        imgs = [np.ascontiguousarray(p, dtype=np.uint8) for p in pages_rgb]
        ws = (C.c_int * n)(*[p.shape[1] for p in imgs])
        hs = (C.c_int * n)(*[p.shape[0] for p in imgs])
        check(self._lib.dsocr_stage_pages(self._h, n, _ptr_array(imgs, C.c_uint8), ws, hs, vs.c()), "stage_pages")
        self._staged_n = n

    def decode_staged(self, seg0: Sequence[int], seg1: Sequence[int], image_token_id: int, params: DecodeParameters
                      ) -> List[DecodeOutcome]:
        n = self._staged_n
        s0 = np.asarray(seg0, dtype=np.int64)
        s1 = np.asarray(seg1, dtype=np.int64)
        outs = [np.zeros(max(1, params.max_new_tokens), dtype=np.int64) for _ in range(n)]
        n_out = (C.c_int * n)()
        n_prompt = (C.c_int * n)()
        p = params.c()
        i64 = C.POINTER(C.c_int64)
        check(self._lib.dsocr_decode_staged(self._h, s0.ctypes.data_as(i64), len(s0), s1.ctypes.data_as(i64), len(s1),
                                            C.c_int64(image_token_id), C.byref(p), C.cast(None, TOKEN_CB), None,
                                            _ptr_array(outs, C.c_int64), n_out, n_prompt), "decode_staged")
        return [DecodeOutcome(n_prompt[i], n_out[i], outs[i][: n_out[i]].tolist()) for i in range(n)]

    def set_stream(self, cuda_stream: int):
        check(self._lib.dsocr_engine_set_stream(self._h, C.c_void_p(cuda_stream)), "set_stream")

    def kernel_timing_begin(self):
        check(self._lib.dsocr_kernel_timing_begin(self._h), "kernel_timing_begin")

    def kernel_timing_end(self) -> list:
        import json

        buf = C.create_string_buffer(1 << 16)
        check(self._lib.dsocr_kernel_timing_end(self._h, buf, len(buf)), "kernel_timing_end")
        return json.loads(buf.value.decode())


def load_model(config_path: str, weights_path: str, snapshot_path: Optional[str] = None, device: int = 0,
               dtype: str = "bf16") -> OcrEngine:
    """`load_model(ModelLoadArgs)` (model/mod.rs:90-115)."""
    return OcrEngine(config_path, weights_path, snapshot_path, device, dtype)
