"""CPU-side checks of the C-ABI library: it loads, exports every symbol the headers declare, its host-only entry
points (integer preprocessing) are bit-exact with the oracle, and it refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import preprocess as P

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    from dsocr.binding import LIB_PATH, lib as load

    if not LIB_PATH.exists():
        ge.build()
    return load()


def _declared_symbols():
    names = []
    for hdr in sorted((ROOT / "include").glob("*.h")):
        text = hdr.read_text()
        names += re.findall(r"DSOCR_API\s+[\w\s\*]+?\b(dsocr_\w+)\s*\(", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_error_plumbing(lib):
    assert b"sm_100a" in lib.dsocr_version()
    lib.dsocr_last_error.restype = C.c_char_p
    assert isinstance(lib.dsocr_last_error(), bytes)


def test_image_token_count(lib):
    assert lib.dsocr_image_token_count(1024, 640, 1, 2, 3) == 903
    assert lib.dsocr_image_token_count(1024, 1024, 0, 1, 1) == 273
    assert lib.dsocr_image_token_count(1024, 640, 1, 1, 1) == 273


@pytest.mark.parametrize("w,h", [(1654, 2339), (700, 500), (333, 517), (2852, 1756), (640, 640), (100, 80), (1000, 3000)])
def test_preprocess_bit_exact_with_oracle(lib, w, h):
    """dsocr_preprocess == vision/resample.rs + vision/preprocess.rs + build_global_view restated by the oracle."""
    from dsocr.engine import VisionSettingsC

    rng = np.random.RandomState(w + h)
    img = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
    g = np.empty((1024, 1024, 3), np.uint8)
    tiles = np.empty((9, 640, 640, 3), np.uint8)
    n, cw, ch = C.c_int(), C.c_int(), C.c_int()
    u8 = C.POINTER(C.c_uint8)
    st = lib.dsocr_preprocess(img.ctypes.data_as(u8), w, h, VisionSettingsC(1024, 640, 1), g.ctypes.data_as(u8),
                              tiles.ctypes.data_as(u8), C.byref(n), C.byref(cw), C.byref(ch))
    assert st == 0
    ref = P.prepare_vision_input(img, 1024, 640, True)
    assert (cw.value, ch.value) == tuple(ref["crop_shape"]) and n.value == len(ref["tiles"])
    assert np.array_equal(g, ref["global"])
    for i, t in enumerate(ref["tiles"]):
        assert np.array_equal(tiles[i], t)


def test_preprocess_mixed_aspect_sweep_covers_every_tile_count(lib):
    """BASELINE configs[2] sweep: aspect ratios that select every tile count n = 2..9 (preprocess.rs:67-138); grid,
    token count, global view and every tile of the C++ host path must equal the oracle bit for bit."""
    from dsocr.engine import VisionSettingsC

    sizes = [(900, 450), (400, 1200), (650, 641), (400, 1600), (2000, 400), (900, 600), (1800, 300), (300, 2100),
             (2400, 300), (2700, 300)]
    seen = set()
    u8 = C.POINTER(C.c_uint8)
    for w, h in sizes:
        rng = np.random.RandomState(w * 7 + h)
        img = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
        g = np.empty((1024, 1024, 3), np.uint8)
        tiles = np.empty((9, 640, 640, 3), np.uint8)
        n, cw, ch = C.c_int(), C.c_int(), C.c_int()
        st = lib.dsocr_preprocess(img.ctypes.data_as(u8), w, h, VisionSettingsC(1024, 640, 1), g.ctypes.data_as(u8),
                                  tiles.ctypes.data_as(u8), C.byref(n), C.byref(cw), C.byref(ch))
        assert st == 0
        ref = P.prepare_vision_input(img, 1024, 640, True)
        assert (cw.value, ch.value) == tuple(ref["crop_shape"]) == P.select_tile_grid(w, h, 640)
        assert n.value == len(ref["tiles"]) == cw.value * ch.value
        assert lib.dsocr_image_token_count(1024, 640, 1, cw.value, ch.value) == P.image_token_count(1024, 640, True, ref["crop_shape"])
        assert np.array_equal(g, ref["global"])
        assert all(np.array_equal(tiles[i], t) for i, t in enumerate(ref["tiles"]))
        seen.add(n.value)
    assert seen == set(range(2, 10))


def test_engine_creation_fails_loudly_without_gpu(lib, tmp_path):
    """No CPU fallback: on a box without an sm_100 device load_model must fail with a clear message."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    st = lib.dsocr_engine_create(b"/nonexistent/config.json", b"/nonexistent/model.safetensors", None, 0, 2, C.byref(h))
    assert st != 0 and not h.value
    lib.dsocr_last_error.restype = C.c_char_p
    msg = lib.dsocr_last_error().decode()
    assert "load_model" in msg and ("CUDA" in msg or "cuda" in msg or "device" in msg), msg


def test_preprocess_property_random_shapes_bit_exact_with_oracle(lib):
    """Seeded random sizes incl. the edges the reference handles specially: 1-pixel sides, sides just below / above the
    640 tile (no-tile rule, preprocess.rs:72-80), extreme aspect ratios (ratio search clamps at 9 tiles), upscales."""
    from dsocr.engine import VisionSettingsC

    rng = np.random.RandomState(2026)
    shapes = [(1, 1), (1, 700), (700, 1), (639, 640), (640, 641), (641, 641), (3000, 90), (90, 2000), (17, 23)]
    shapes += [(int(rng.randint(2, 1500)), int(rng.randint(2, 1500))) for _ in range(12)]
    u8 = C.POINTER(C.c_uint8)
    for w, h in shapes:
        img = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
        for base, tile, crop in ((1024, 640, 1), (640, 640, 0)):
            gsz = base if crop else tile
            g = np.empty((gsz, gsz, 3), np.uint8)
            tiles = np.empty((9, tile, tile, 3), np.uint8)
            n, cw, ch = C.c_int(), C.c_int(), C.c_int()
            st = lib.dsocr_preprocess(img.ctypes.data_as(u8), w, h, VisionSettingsC(base, tile, crop), g.ctypes.data_as(u8),
                                      tiles.ctypes.data_as(u8), C.byref(n), C.byref(cw), C.byref(ch))
            assert st == 0, (w, h, lib.dsocr_last_error().decode())
            ref = P.prepare_vision_input(img, base, tile, bool(crop))
            assert (cw.value, ch.value) == tuple(ref["crop_shape"] or (1, 1)) and n.value == len(ref["tiles"]), (w, h, crop)
            assert np.array_equal(g, ref["global"]), (w, h, crop)
            for i, t in enumerate(ref["tiles"]):
                assert np.array_equal(tiles[i], t), (w, h, i)


def test_plain_c_consumer_builds_links_and_runs(lib, tmp_path):
    """examples/c_consumer.c: the boundary is usable from C99 with nothing but include/dsocr.h and -ldsocr (what a cgo /
    Rust FFI binding needs); its host-only calls run here, engine creation fails loudly without a GPU."""
    import shutil
    import subprocess

    from dsocr.binding import LIB_PATH

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = tmp_path / "c_consumer"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "examples" / "c_consumer.c"), "-L",
           str(LIB_PATH.parent), "-ldsocr", f"-Wl,-rpath,{LIB_PATH.parent}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(tmp_path / "snap")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "snapshot: 1 tensor(s), `linear.weight` [2, 32] dtype 8, 68 payload bytes, 8 bias bytes" in r.stdout
    assert "A4 page: grid 2x3, 6 tiles, 903 image tokens" in r.stdout
    import torch
    if not torch.cuda.is_available():
        assert "engine_create failed as expected" in r.stdout
