#!/usr/bin/env python
"""Writes tests/golden/fixture_tiny/: one synthetic fixture page in the REFERENCE's baseline schema
(long_generation_baseline.rs:30-150) produced by the CPU oracle on the tiny random-init model (tests/helpers.tiny_model,
seed 1234, bf16 storage).  No tokenizer exists offline, so `prompt` / `rendered_prompt` are descriptive and the token ids
of the text part are the fixed ids below."""
import json
import sys
from pathlib import Path

import numpy as np
import torch
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import decoder as D, preprocess as P, vision as V  # noqa: E402
from tests.helpers import tiny_model  # noqa: E402

OUT = Path(__file__).resolve().parent / "fixture_tiny"
TAIL = [201, 202, 203, 204]
MAX_NEW = 24


def main():
    OUT.mkdir(exist_ok=True)
    cfg, ck, _ = tiny_model("bf16")
    page = P.synthetic_page(640, 480, seed=21)
    Image.fromarray(page).save(OUT / "page.png")
    page = np.asarray(Image.open(OUT / "page.png").convert("RGB"))
    vi = P.prepare_vision_input(page, 640, 640, False)
    with torch.no_grad():
        rows = V.VisionOracle(cfg, ck).encode(torch.from_numpy(P.image_to_tensor(vi["global"])), None, None)
        ids, mask = D.build_prompt_tokens([[], TAIL], [rows.shape[0]], cfg)
        gen = D.DecoderOracle(cfg, ck).generate(ids, mask, rows, MAX_NEW, 20, cfg.eos_token_id if hasattr(cfg, "eos_token_id") else 1)
    start = mask.index(1)
    (OUT / "baseline.json").write_text(json.dumps({
        "variant": "ocr1", "prompt": "<image>\n(text ids 201..204; no tokenizer offline)", "image": "page.png", "base_size": 640,
        "image_size": 640, "crop_mode": False, "max_new_tokens": MAX_NEW,
        "model": "tiny random-init DeepSeek-OCR (tests/helpers.tiny_model, seed 1234, bf16 storage)"}, indent=1) + "\n")
    (OUT / "prompt.json").write_text(json.dumps({
        "rendered_prompt": "<image>\n(text ids 201..204; no tokenizer offline)", "input_ids": ids, "images_seq_mask": mask,
        "image_token_ranges": [{"start": start, "length": int(sum(mask))}], "image_token_counts": [int(sum(mask))],
        "vision_token_counts": [int(rows.shape[0])], "vision_token_total": int(rows.shape[0]), "bos_token_id": 0,
        "image_token_id": cfg.image_token_id, "prefill_len": len(ids)}) + "\n")
    (OUT / "output_tokens.json").write_text(json.dumps({
        "tokens": ids + gen, "prefill_len": len(ids), "generated_len": MAX_NEW, "eos_token_id": 1}) + "\n")
    np.savez_compressed(OUT / "fused_tokens.npz", fused_tokens_image0=rows.numpy().astype(np.float32))
    print("wrote", OUT, "generated", gen)


if __name__ == "__main__":
    main()
