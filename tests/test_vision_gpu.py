"""Vision path parity (SAM -> CLIP -> projector -> token layout) through the C ABI against the f32 oracle
that restates vision/sam.rs, vision/clip.rs and model/mod.rs:392-444,590-923, at the reference's own tap
points (SamDebugTrace / ClipDebugTrace / VisionProjectionOutputs).

Tolerances (bf16 operands, f32 accumulate/residual/norm/softmax vs an f32-math oracle on the same
bf16-rounded weights): max-abs <= 3% of the tap's max magnitude and cosine >= 0.999.  For context the
reference's own gates against the HF model are max-abs 5.0 (pre-projector) and 2.0 (fused tokens):
crates/infer-deepseek/tests/baseline.rs:335, :805."""
import numpy as np
import pytest
import torch

from oracle import preprocess as P
from oracle import vision as V
from tests.helpers import report, tiny_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from dsocr.engine import load_model

    cfg, ck, d = tiny_model("bf16")
    eng = load_model(d + "/config.json", d + "/model.safetensors", None, 0, "bf16")
    eng.set_option("record_taps", 1)
    yield cfg, ck, eng, V.VisionOracle(cfg, ck)
    eng.close()


def _check(name, got, ref, rel=0.03, min_cos=0.999):
    err, scale, c = report(name, got, ref)
    assert err <= rel * scale, (name, err, scale)
    assert c >= min_cos, (name, c)


@pytest.mark.parametrize("size", [640, 512, 1024])
def test_base_mode_taps_and_rows(setup, size):
    """Base-style mode (crop_mode=false): one global view."""
    cfg, ck, eng, oracle = setup
    page = P.synthetic_page(size + 37, size - 11, seed=size)
    vi = P.prepare_vision_input(page, size, size, False)
    g = P.image_to_tensor(vi["global"])
    rows = torch.from_numpy(eng.vision_encode(g, None, None))
    trace_s, trace_c, taps = {}, {}, {}
    gt = torch.from_numpy(g).unsqueeze(0)
    sam = oracle.sam.forward(gt, trace_s)
    clip = oracle.clip.forward(sam, trace_c)
    ref_rows = oracle.encode(gt, None, None, taps)
    assert rows.shape == ref_rows.shape == (P.image_token_count(size, size, False, None), cfg.n_embed)
    tok = (size // 16) ** 2
    _check("sam.pos_added", torch.from_numpy(eng.tap("sam.pos_added")).reshape(tok, -1), trace_s["pos_added"].reshape(tok, -1), 0.01)
    for i in range(cfg.sam_depth):
        _check(f"sam.block.{i}", torch.from_numpy(eng.tap(f"sam.block.{i}")).reshape(tok, -1),
               trace_s["block_outputs"][i].reshape(tok, -1))
    _check("sam.neck_conv1", torch.from_numpy(eng.tap("sam.neck_conv1")).reshape(tok, -1),
           trace_s["neck_conv1"][0].permute(1, 2, 0).reshape(tok, -1))
    _check("sam.neck_conv2", torch.from_numpy(eng.tap("sam.neck_conv2")).reshape(tok, -1),
           trace_s["neck_conv2"][0].permute(1, 2, 0).reshape(tok, -1))
    n = (size // 64) ** 2
    _check("sam.net3", torch.from_numpy(eng.tap("sam.net3")).reshape(n, -1), sam[0].permute(1, 2, 0).reshape(n, -1))
    _check("clip.embeddings", torch.from_numpy(eng.tap("clip.embeddings")).reshape(n + 1, -1), trace_c["embeddings"][0])
    for i in range(cfg.clip_layers):
        _check(f"clip.layer.{i}", torch.from_numpy(eng.tap(f"clip.layer.{i}")).reshape(n + 1, -1), trace_c["layer_outputs"][i][0])
    _check("global_pre", torch.from_numpy(eng.tap("global_pre")).reshape(n, -1), taps["global_pre"][0])
    _check("global_post", torch.from_numpy(eng.tap("global_post")).reshape(n, -1), taps["global_post"][0])
    _check("fused_tokens", rows, ref_rows)
    # structural rows are exact copies of the learned newline / separator vectors
    q = size // 64
    assert torch.equal(rows[q], ck["model.image_newline"].float())
    assert torch.equal(rows[-1], ck["model.view_seperator"].float())


def test_gundam_mode_rows(setup):
    """crop_mode=true: n x 640 local crops + 1024 global view; token layout [local ; global ; separator]."""
    cfg, ck, eng, oracle = setup
    page = P.synthetic_page(700, 1400, seed=7)
    vi = P.prepare_vision_input(page, 1024, 640, True)
    assert vi["crop_shape"] == (1, 2) and len(vi["tiles"]) == 2
    g = P.image_to_tensor(vi["global"])
    tiles = np.stack([P.image_to_tensor(t) for t in vi["tiles"]])
    rows = torch.from_numpy(eng.vision_encode(g, tiles, vi["crop_shape"]))
    ref = oracle.encode(torch.from_numpy(g), torch.from_numpy(tiles), vi["crop_shape"])
    assert rows.shape == ref.shape == (P.image_token_count(1024, 640, True, vi["crop_shape"]), cfg.n_embed)
    n_local = rows.shape[0] - (16 * 17 + 1)
    report("gundam local rows", rows[:n_local], ref[:n_local])
    report("gundam global rows", rows[n_local:], ref[n_local:])
    _check("gundam fused_tokens", rows, ref)


def test_u8_batch_matches_f32_entry(setup):
    """The fused u8 normalise+patchify entry gives the same rows as the reference-shaped f32 CHW entry."""
    cfg, ck, eng, oracle = setup
    pages = [P.synthetic_page(640, 640, seed=s) for s in (1, 2, 3)]
    outs = eng.vision_encode_u8(pages, [None] * 3, [(1, 1)] * 3, 640)
    for pg, o in zip(pages, outs):
        single = eng.vision_encode(P.image_to_tensor(pg), None, None)
        assert np.array_equal(o, single)


def test_preprocess_bit_exact_with_oracle(setup):
    """dsocr_preprocess (C++ integer resampler / tiler) == oracle restatement of vision/resample.rs."""
    from dsocr.engine import VisionSettings

    cfg, ck, eng, oracle = setup
    rng = np.random.RandomState(0)
    for (w, h) in [(1654, 2339), (700, 500), (333, 517), (2852, 1756), (640, 640), (100, 80)]:
        img = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
        g, tiles, crop = eng.preprocess(img, VisionSettings(1024, 640, True))
        ref = P.prepare_vision_input(img, 1024, 640, True)
        assert crop == tuple(ref["crop_shape"])
        assert np.array_equal(g, ref["global"])
        assert tiles.shape[0] == len(ref["tiles"])
        for a, b in zip(tiles, ref["tiles"]):
            assert np.array_equal(a, b)


def test_gpu_preprocess_bit_exact_with_oracle(setup):
    """Device integer resampler / tiler (resample_h / resample_v kernels) == oracle restatement of vision/resample.rs,
    including an upscale (where resample.rs deliberately differs from Pillow) and the no-crop small-image case."""
    from dsocr.engine import VisionSettings

    cfg, ck, eng, oracle = setup
    rng = np.random.RandomState(1)
    # the second half of the list is the mixed-aspect sweep of BASELINE configs[2]: tile counts 2..9
    for (w, h) in [(1654, 2339), (700, 500), (333, 517), (2852, 1756), (640, 640), (100, 80), (1024, 1024),
                   (900, 450), (400, 1200), (400, 1600), (2000, 400), (1800, 300), (300, 2100), (2400, 300), (2700, 300)]:
        img = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
        g, tiles, crop = eng.preprocess_gpu(img, VisionSettings(1024, 640, True))
        ref = P.prepare_vision_input(img, 1024, 640, True)
        assert crop == tuple(ref["crop_shape"]), (w, h)
        assert np.array_equal(g, ref["global"]), (w, h)
        assert tiles.shape[0] == len(ref["tiles"])
        for a, b in zip(tiles, ref["tiles"]):
            assert np.array_equal(a, b), (w, h)
