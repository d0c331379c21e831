"""Reference-schema baseline fixtures (dsocr/baselines.py; schema of
crates/infer-deepseek/tests/long_generation_baseline.rs:30-150): the committed synthetic fixture loads with the
reference test's consistency checks, the CPU oracle reproduces its tokens and vision rows, and `run_baseline` drives an
engine from the fixture's own ids (stub engine here; scripts/run_baseline.py does it on a GPU)."""
import json
import os
import shutil

import numpy as np
import pytest
import torch

from dsocr import baselines
from oracle import decoder as D, preprocess as P, vision as V
from tests.helpers import tiny_model

FIX = os.path.join(os.path.dirname(__file__), "golden", "fixture_tiny")


def test_fixture_loads_with_the_reference_checks():
    b = baselines.load_baseline(FIX)
    assert b.variant == "ocr1" and (b.base_size, b.image_size, b.crop_mode) == (640, 640, False)
    seg0, seg1, n_img = b.segments()
    assert seg0 == [] and seg1 == [201, 202, 203, 204] and n_img == b.vision_token_total == 111
    assert b.input_ids[0] == 0 and len(b.expected) == b.requested_tokens == 24
    assert [0] + seg0 + [b.image_token_id] * n_img + seg1 == b.input_ids


def test_loader_rejects_inconsistent_fixtures(tmp_path):
    d = tmp_path / "bad"
    shutil.copytree(FIX, d)
    out = json.loads((d / "output_tokens.json").read_text())
    out["prefill_len"] += 1
    (d / "output_tokens.json").write_text(json.dumps(out))
    with pytest.raises(ValueError, match="output prefill_len"):
        baselines.load_baseline(str(d))
    out["prefill_len"] -= 1
    out["tokens"].append(1)  # trailing EOS is dropped from the expectation (expected_generated_tokens)
    (d / "output_tokens.json").write_text(json.dumps(out))
    assert len(baselines.load_baseline(str(d)).expected) == 24


def test_oracle_reproduces_the_fixture():
    from PIL import Image

    b = baselines.load_baseline(FIX)
    cfg, ck, _ = tiny_model("bf16")
    page = np.asarray(Image.open(b.image).convert("RGB"))
    vi = P.prepare_vision_input(page, b.base_size, b.image_size, b.crop_mode)
    with torch.no_grad():
        rows = V.VisionOracle(cfg, ck).encode(torch.from_numpy(P.image_to_tensor(vi["global"])), None, None)
        ref_rows = torch.from_numpy(np.load(os.path.join(FIX, "fused_tokens.npz"))["fused_tokens_image0"])
        assert rows.shape == ref_rows.shape and (rows - ref_rows).abs().max().item() <= 1e-4 * ref_rows.abs().max().item()
        gen = D.DecoderOracle(cfg, ck).generate(b.input_ids, b.images_seq_mask, rows, b.requested_tokens, 20, b.eos_token_id)
    res = baselines.compare(b.expected, gen)
    assert res["match"] and res["agreement"] == 1.0, res


def test_run_baseline_drives_the_engine_from_fixture_ids():
    from dsocr.engine import DecodeOutcome

    b = baselines.load_baseline(FIX)
    seen = {}

    class Stub:
        def decode_pages(self, pages, vs, seg0, seg1, image_id, params):
            seen.update(shape=pages[0].shape, vs=(vs.base_size, vs.image_size, vs.crop_mode), seg0=list(seg0), seg1=list(seg1),
                        image_id=image_id, max_new=params.max_new_tokens, ngram=params.no_repeat_ngram_size, eos=params.eos_token_id)
            got = list(b.expected)
            got[5] = got[5] + 1
            return [DecodeOutcome(len(b.input_ids), len(got), got)]

    res = baselines.run_baseline(Stub(), b)
    assert seen == {"shape": (480, 640, 3), "vs": (640, 640, False), "seg0": [], "seg1": [201, 202, 203, 204],
                    "image_id": b.image_token_id, "max_new": 24, "ngram": 20, "eos": 1}
    assert res["match"] is False and res["first_mismatch"] == 5 and abs(res["agreement"] - 23 / 24) < 1e-9
    assert res["prompt_tokens"] == res["prompt_tokens_expected"] == 116
