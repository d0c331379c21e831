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
