"""Parity at the FULL architecture (12 SAM blocks, 24 CLIP layers, 12 decoder layers, 64 routed experts, vocabulary
129 280) with random-init weights of the exact shapes: the instantiations the tiny fixtures never reach
(`post_attn_kernel<T,64>`, `router_kernel<64>`, `dsq_router_kernel<64,16>`, the 129 280-wide lm_head + select_token) and
the error accumulated over the real depth.  Same tolerances as the tiny-model tests."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder as D
from oracle import vision as V
from tests.helpers import full_model, report

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from dsocr.engine import load_model

    torch.set_num_threads(max(1, torch.get_num_threads()))
    cfg, ck, d = full_model("bf16")
    eng = load_model(d + "/config.json", d + "/model.safetensors", None, 0, "bf16")
    yield cfg, ck, eng
    eng.close()


@pytest.mark.parametrize("size", [640, 1024])
def test_full_depth_vision_taps(setup, size):
    """Every SamDebugTrace / ClipDebugTrace tap of all 12 + 24 layers and the projected rows, at both view sizes."""
    from tests.test_vision_gpu import test_base_mode_taps_and_rows as taps_check

    cfg, ck, eng = setup
    eng.set_option("record_taps", 1)
    try:
        with torch.no_grad():
            taps_check((cfg, ck, eng, V.VisionOracle(cfg, ck)), size)
    finally:
        eng.set_option("record_taps", 0)


def _prompts(cfg, n_imgs, seed):
    g = torch.Generator().manual_seed(seed)
    ids, masks, rows = [], [], []
    for n_img in n_imgs:
        text = torch.randint(2, 100000, (6 + n_img % 5,), generator=g).tolist()
        t, m = D.build_prompt_tokens([[], text] if n_img else [text], [n_img] if n_img else [], cfg)
        ids.append(t); masks.append(m)
        rows.append((torch.randn(n_img, cfg.hidden_size, generator=g) * 0.7).numpy() if n_img else None)
    return ids, masks, rows


def test_full_decoder_batched_and_batch1(setup):
    """5 pages (Gundam-, Base- and text-sized prompts) through the batched decode step for 64 tokens: teacher-forced logits
    and free-running tokens against the oracle with the f32 cache, token agreement with the f16 cache; then one page
    through the fused small-batch step."""
    from dsocr.engine import DecodeParameters

    cfg, ck, eng = setup
    oracle = D.DecoderOracle(cfg, ck)
    steps = 64
    ids, masks, rows = _prompts(cfg, [903, 273, 100, 30, 0], seed=3)
    g = torch.Generator().manual_seed(4)
    forced = [torch.randint(2, 100000, (steps,), generator=g).tolist() for _ in ids]
    params = DecodeParameters(max_new_tokens=steps, no_repeat_ngram_size=20, eos_token_id=None)
    sel, logits = eng.generate_forced(ids, masks, rows, params, forced, want_logits=True)
    free = eng.generate_batch(ids, masks, rows, params)
    # post_attn_kernel<T, 64, R> with R = 2 / 4 / 8 rows per block (picked for large steps) forced onto these 5 pages: the
    # per-row arithmetic does not depend on the rows-per-block choice, so the tokens must be the same ones
    for r in ("2", "4", "8"):
        os.environ["DSOCR_POST_ATTN_ROWS"] = r
        try:
            free_r = eng.generate_batch(ids, masks, rows, params)
        finally:
            del os.environ["DSOCR_POST_ATTN_ROWS"]
        assert free_r == free, f"{r} rows per block"
    eng.set_option("kv_cache_f16", 1)
    try:
        free16 = eng.generate_batch(ids, masks, rows, params)
    finally:
        eng.set_option("kv_cache_f16", 0)
    agree16 = n = 0
    with torch.no_grad():
        for p in range(len(ids)):
            rt = None if rows[p] is None else torch.from_numpy(rows[p])
            lg = []
            ref_sel = oracle.generate(ids[p], masks[p], rt, steps, 20, None, forced=forced[p], logits_out=lg)
            ref_free = oracle.generate(ids[p], masks[p], rt, steps, 20, None)
            err, scale, c = report(f"full arch, page {p} (prompt {len(ids[p])}), teacher-forced logits", torch.from_numpy(logits[p]), torch.stack(lg))
            assert err <= 2e-3 * scale and c > 0.99999
            assert sel[p] == ref_sel
            assert free[p] == ref_free
            agree16 += sum(int(a == b) for a, b in zip(free16[p], ref_free)); n += steps
            if p == 1:  # the same page alone: <= 4 pages take the fused small-batch step
                one = eng.generate_batch([ids[p]], [masks[p]], [rows[p]], params)[0]
                assert one == ref_free
    print(f"[parity] full arch, f16 KV: free-running agreement {agree16 / n:.4f} over {n} tokens")
    assert agree16 / n >= 0.95
