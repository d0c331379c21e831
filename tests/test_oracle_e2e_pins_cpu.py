"""The oracle's towers held to outputs of INDEPENDENT implementations (tests/golden/make_golden_towers.py):
vLLM's PyTorch SAM ViT (deepencoder.py), Hugging Face's CLIPEncoder + vLLM's pos-embed resize, and Hugging Face's
LlamaForCausalLM with transformers' DeepseekV2Moe blocks, all loaded with the same seeded tiny checkpoint.  The stored
vectors are those implementations' outputs; the oracle must reproduce them to f32 rounding (<= 1e-4 absolute on O(5)
values; measured 2e-5 / 1.4e-5 / 1.6e-6).  This is what ties oracle/vision.py and oracle/decoder.py - and through them
every CUDA parity test - to the model the reference implements, since the reference itself (Rust + candle) cannot be
built or run offline."""
import hashlib
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import config as OC, decoder as D, vision as V

GOLD = Path(__file__).resolve().parent / "golden" / "towers_tiny.npz"


@pytest.fixture(scope="module")
def setup():
    import importlib.util
    import sys

    spec = importlib.util.spec_from_file_location("make_golden_towers", GOLD.parent / "make_golden_towers.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    cfg = OC.tiny_config()
    ck = OC.random_checkpoint(cfg, seed=1234, storage=torch.bfloat16)
    gold = np.load(GOLD)
    if gen.checkpoint_digest(ck) != bytes(gold["digest"]).decode():
        pytest.skip("torch's CPU generator produced a different seeded checkpoint than the one the vectors were made with")
    return gen, cfg, ck, gold


@pytest.mark.parametrize("size", [640, 1024])
def test_sam_and_clip_match_vllm_and_hf(setup, size):
    gen, cfg, ck, gold = setup
    x = gen.view(size, seed=size)
    with torch.no_grad():
        sam = V.SamOracle(cfg, ck).forward(x)
        clip = V.ClipOracle(cfg, ck).forward(sam)
    es = np.abs(sam[0, ::4].numpy() - gold[f"sam_{size}"]).max()
    ec = np.abs(clip[0, :, ::4].numpy() - gold[f"clip_{size}"]).max()
    print(f"[pin] SAM {size}: max-abs vs vLLM {es:.3e}; CLIP: max-abs vs HF {ec:.3e}")
    assert gold[f"sam_{size}"].shape == (256, size // 64, size // 64)
    assert es <= 1e-4 and ec <= 1e-4


def test_decoder_matches_hf_llama_deepseek_moe(setup):
    gen, cfg, ck, gold = setup
    ids, mask, rows, forced = gen.decoder_case(cfg)
    lg = []
    with torch.no_grad():
        D.DecoderOracle(cfg, ck).generate(ids, mask, rows, len(forced), 20, None, forced=forced, logits_out=lg)
    got = torch.stack(lg).numpy()
    err = np.abs(got - gold["dec_logits"]).max()
    print(f"[pin] decoder logits (prefill + {len(forced) - 1} cached steps): max-abs vs HF {err:.3e}")
    assert err <= 1e-4
    assert (got.argmax(-1) == gold["dec_logits"].argmax(-1)).all()
