"""DSQ path parity: engine loaded with a synthesised q8_0 / q4k / q6k snapshot vs the f32 oracle run on the SAME
dequantised weights (contract of SURVEY.md 8c 'Parity note for DSQ': y = x . dequant(W)^T in f32; candle's own CPU and
CUDA backends differ from each other because they re-quantise activations)."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder as D
from oracle import dsq
from tests.helpers import report, tiny_model
from tests.test_decoder_gpu import _prompts

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("primary,name", [(dsq.Q8_0, "q8_0"), (dsq.Q4K, "q4k"), (dsq.Q6K, "q6k")])
def test_dsq_decode_matches_dequantised_oracle(primary, name):
    from dsocr.engine import DecodeParameters, load_model

    cfg, ck, d = tiny_model("bf16")
    snap = os.path.join(d, f"model.{name}.dsq")
    assigned = dsq.write_model_snapshot(snap, cfg, ck, primary)
    # dtype assignment follows the exporter: in_dim 1280 / 1792 -> primary, 896 / 6848 -> Q8_0 fallback, lm_head Q8_0
    assert assigned["lm_head.weight"] == dsq.Q8_0
    assert assigned["model.layers.1.mlp.experts.0.down_proj.weight"] == dsq.Q8_0
    assert assigned["model.layers.1.mlp.experts.0.gate_proj.weight"] == primary
    oracle = D.DecoderOracle(cfg, dsq.dequantized_checkpoint(snap, ck))
    eng = load_model(d + "/config.json", d + "/model.safetensors", snap, 0, "bf16")
    assert eng.info.quantized == 1
    ids, masks, rows = _prompts(cfg, [21, 0, 40], seed=17)
    steps = 10
    g = torch.Generator().manual_seed(4)
    forced = [torch.randint(2, cfg.vocab_size - 2, (steps,), generator=g).tolist() for _ in ids]
    params = DecodeParameters(max_new_tokens=steps, eos_token_id=None)
    sel, logits = eng.generate_forced(ids, masks, rows, params, forced, want_logits=True)
    for p in range(len(ids)):
        ref_logits = []
        ref_sel = oracle.generate(ids[p], masks[p], None if rows[p] is None else torch.from_numpy(rows[p]), steps, 20, None,
                                  forced=forced[p], logits_out=ref_logits)
        err, scale, c = report(f"dsq {name} teacher-forced logits page {p}", torch.from_numpy(logits[p]), torch.stack(ref_logits))
        assert err <= 2e-3 * scale and c > 0.99999
        assert sel[p] == ref_sel
    free = eng.generate_batch(ids, masks, rows, DecodeParameters(max_new_tokens=16, eos_token_id=None))
    for p in range(len(ids)):
        ref = oracle.generate(ids[p], masks[p], None if rows[p] is None else torch.from_numpy(rows[p]), 16, 20, None)
        assert free[p] == ref
    eng.close()


@pytest.mark.parametrize("kv_f16", [0, 1])
def test_dsq_batch1_fused_step_long_context(kv_f16):
    """Batch-1 decode (the fused 6-launches-per-layer step of dsq_decode.cu, MT = 1 kernels) over enough steps that
    the split-key attention runs several key ranges per (row, head): teacher-forced logits vs the f32 oracle on the
    same dequantised weights, then the free-running tokens."""
    from dsocr.engine import DecodeParameters, load_model

    cfg, ck, d = tiny_model("bf16")
    snap = os.path.join(d, "model.q4k_b1.dsq")
    dsq.write_model_snapshot(snap, cfg, ck, dsq.Q4K)
    oracle = D.DecoderOracle(cfg, dsq.dequantized_checkpoint(snap, ck))
    eng = load_model(d + "/config.json", d + "/model.safetensors", snap, 0, "bf16")
    eng.set_option("kv_cache_f16", kv_f16)
    ids, masks, rows = _prompts(cfg, [33], seed=5)
    steps = 150
    g = torch.Generator().manual_seed(9)
    forced = [torch.randint(2, cfg.vocab_size - 2, (steps,), generator=g).tolist()]
    params = DecodeParameters(max_new_tokens=steps, eos_token_id=None)
    sel, logits = eng.generate_forced(ids, masks, rows, params, forced, want_logits=True)
    ref_logits = []
    ref_sel = oracle.generate(ids[0], masks[0], torch.from_numpy(rows[0]), steps, 20, None, forced=forced[0], logits_out=ref_logits)
    err, scale, c = report(f"dsq q4k batch-1 fused step, {steps} forced steps, kv_f16={kv_f16}", torch.from_numpy(logits[0]), torch.stack(ref_logits))
    assert err <= (4e-3 if kv_f16 else 2e-3) * scale and c > 0.9999
    if not kv_f16:
        assert sel[0] == ref_sel
        free = eng.generate_batch(ids, masks, rows, DecodeParameters(max_new_tokens=48, eos_token_id=None))
        ref = oracle.generate(ids[0], masks[0], torch.from_numpy(rows[0]), 48, 20, None)
        assert free[0] == ref
    eng.close()


def test_dsq_fused_step_matches_unfused_path(monkeypatch):
    """The fused step and the per-linear GEMV path (DSOCR_DSQ_UNFUSED=1) are two schedules of the same f32 math."""
    from dsocr.engine import DecodeParameters, load_model

    cfg, ck, d = tiny_model("bf16")
    snap = os.path.join(d, "model.q6k_ab.dsq")
    dsq.write_model_snapshot(snap, cfg, ck, dsq.Q6K)
    ids, masks, rows = _prompts(cfg, [12, 30], seed=23)
    steps = 12
    g = torch.Generator().manual_seed(2)
    forced = [torch.randint(2, cfg.vocab_size - 2, (steps,), generator=g).tolist() for _ in ids]
    params = DecodeParameters(max_new_tokens=steps, eos_token_id=None)
    out = []
    for unfused in (False, True):
        if unfused:
            monkeypatch.setenv("DSOCR_DSQ_UNFUSED", "1")
        eng = load_model(d + "/config.json", d + "/model.safetensors", snap, 0, "bf16")
        out.append(eng.generate_forced(ids, masks, rows, params, forced, want_logits=True)[1])
        eng.close()
    for p in range(len(ids)):
        err, scale, c = report(f"dsq q6k fused vs unfused page {p}", torch.from_numpy(out[0][p]), torch.from_numpy(out[1][p]))
        assert err <= 1e-4 * scale


@pytest.mark.parametrize("primary,name", [(dsq.Q4K, "q4k"), (dsq.Q8_0, "q8_0")])
def test_dsq_gemm_prefill_and_batched_decode(primary, name):
    """Prefill (> 256 prompt rows) and decode steps of > 4 pages of a DSQ engine run the dequant-fused tensor-core GEMM
    (csrc/linear_dq.cuh) through the same decoder schedule as the float engine: teacher-forced logits and free-running
    tokens against the f32 oracle on the same dequantised weights, and timing against the per-row GEMV path it replaces."""
    import time

    from dsocr.engine import DecodeParameters, load_model

    cfg, ck, d = tiny_model("bf16")
    snap = os.path.join(d, f"model.{name}_gemm.dsq")
    dsq.write_model_snapshot(snap, cfg, ck, primary)
    oracle = D.DecoderOracle(cfg, dsq.dequantized_checkpoint(snap, ck))
    eng = load_model(d + "/config.json", d + "/model.safetensors", snap, 0, "bf16")
    ids, masks, rows = _prompts(cfg, [273, 40, 0, 130, 7, 64], seed=29)
    steps = 20
    g = torch.Generator().manual_seed(6)
    forced = [torch.randint(2, cfg.vocab_size - 2, (steps,), generator=g).tolist() for _ in ids]
    params = DecodeParameters(max_new_tokens=steps, eos_token_id=None)
    sel, logits = eng.generate_forced(ids, masks, rows, params, forced, want_logits=True)
    t0 = time.perf_counter()
    free = eng.generate_batch(ids, masks, rows, params)
    t_gemm = time.perf_counter() - t0
    for p in range(len(ids)):
        rt = None if rows[p] is None else torch.from_numpy(rows[p])
        ref_logits = []
        ref_sel = oracle.generate(ids[p], masks[p], rt, steps, 20, None, forced=forced[p], logits_out=ref_logits)
        err, scale, c = report(f"dsq {name} GEMM path, page {p} (prompt {len(ids[p])})", torch.from_numpy(logits[p]), torch.stack(ref_logits))
        assert err <= 2e-3 * scale and c > 0.99999
        assert sel[p] == ref_sel
        assert free[p] == oracle.generate(ids[p], masks[p], rt, steps, 20, None)
    eng.close()
    os.environ["DSOCR_DSQ_GEMV"] = "1"
    try:
        eng = load_model(d + "/config.json", d + "/model.safetensors", snap, 0, "bf16")
        eng.generate_batch(ids, masks, rows, DecodeParameters(max_new_tokens=2, eos_token_id=None))
        t0 = time.perf_counter()
        free_gemv = eng.generate_batch(ids, masks, rows, params)
        t_gemv = time.perf_counter() - t0
        eng.close()
    finally:
        del os.environ["DSOCR_DSQ_GEMV"]
    assert free_gemv == free
    print(f"[timing] dsq {name} 6 pages / {sum(len(i) for i in ids)} prompt rows / {steps} steps: GEMM path {t_gemm * 1e3:.1f} ms, "
          f"per-row GEMV path {t_gemv * 1e3:.1f} ms")
