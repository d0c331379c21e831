"""SURVEY 8 (f1)/(f2) against the REAL engine on the GPU (round 1 only had stub-engine tests): the CLI-equivalent driver
writes the reference's --output-json / --bench-output files from an actual decode, and the OpenAI-compatible serving loop
(batcher -> dsocr_decode_requests) answers concurrent JSON and SSE requests - including a two-image prompt and a seeded
sampling request - with the tokens the engine produces for the same requests directly."""
import base64
import importlib.util
import io
import json
import threading
from pathlib import Path

import numpy as np
import pytest

from oracle import preprocess as P
from tests.helpers import tiny_model

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


class Tok:
    """Deterministic stand-in for tokenizers.Tokenizer (the tokenizer stays with the host language)."""

    def encode(self, text, add_special_tokens=False):
        return type("Enc", (), {"ids": [100 + (sum(map(ord, w)) % 900) for w in text.split()]})()

    def token_to_id(self, t):
        return 2047 if t == "<image>" else None

    def decode(self, ids, skip_special_tokens=False):
        return " ".join(f"t{i}" for i in ids)


def _png(page: np.ndarray, path=None):
    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(page).save(path or buf, format="PNG")
    return None if path else "data:image/png;base64," + base64.b64encode(buf.getvalue()).decode()


def test_cli_driver_end_to_end(tmp_path, monkeypatch):
    import tokenizers

    from dsocr.engine import DecodeParameters, VisionSettings, load_model

    spec = importlib.util.spec_from_file_location("dsocr_cli", ROOT / "scripts" / "dsocr_cli.py")
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    monkeypatch.setattr(tokenizers.Tokenizer, "from_file", staticmethod(lambda path: Tok()))
    cfg, ck, d = tiny_model("bf16")
    page = P.synthetic_page(700, 900, seed=3)
    _png(page, tmp_path / "page.png")
    out_json, bench_json = tmp_path / "rust_output.json", tmp_path / "bench_raw.json"
    rc = cli.main(["--model", "deepseek-ocr", "--image", str(tmp_path / "page.png"), "--device", "cuda:0", "--dtype", "bf16",
                   "--max-new-tokens", "16", "--bench", "--bench-output", str(bench_json), "--output-json", str(out_json),
                   "--prompt", "<image>\nFree OCR.", "--model-config", d + "/config.json", "--weights", d + "/model.safetensors",
                   "--tokenizer", "tok.json", "--eos-token-id", "-1", "--quiet"])
    assert rc == 0
    o = json.loads(out_json.read_text())
    eng = load_model(d + "/config.json", d + "/model.safetensors", None, 0, "bf16")
    seg1 = Tok().encode("Free OCR.").ids
    ref = eng.decode_pages([page], VisionSettings(1024, 640, True), [], seg1, 2047, DecodeParameters(16, eos_token_id=None))[0]
    eng.close()
    assert o["tokens"] == ref.generated_tokens and o["prompt_tokens"] == ref.prompt_tokens and o["generated_len"] == 16
    assert o["decoded"] == Tok().decode(ref.generated_tokens)
    stages = {s["stage"]: s["total_ms"] for s in json.loads(bench_json.read_text())["stage_totals"]}
    assert stages["vision.compute_embeddings"] > 0 and stages["decode.prefill"] > 0 and stages["decode.iterative"] > 0


def test_server_against_engine():
    from fastapi.testclient import TestClient

    from dsocr.batcher import PageBatcher, engine_runner
    from dsocr.engine import DecodeParameters, VisionSettings, load_model
    from dsocr.server import create_app, params_from_tuple

    cfg, ck, d = tiny_model("bf16")
    eng = load_model(d + "/config.json", d + "/model.safetensors", None, 0, "bf16")
    batcher = PageBatcher(engine_runner(eng, params_from_tuple, lambda v: VisionSettings(*v)), max_batch=8, max_wait_ms=200)
    tok = Tok()
    app = create_app(batcher, tok, 2047, vision=(1024, 640, True), max_new_tokens=12)
    client = TestClient(app)
    pages = [P.synthetic_page(700, 900, seed=1), P.synthetic_page(900, 500, seed=2), P.synthetic_page(640, 640, seed=3)]

    def body(content, **kw):
        return {"model": "deepseek-ocr", "messages": [{"role": "user", "content": content}], "max_tokens": 12, **kw}

    img = lambda p: {"type": "image_url", "image_url": {"url": _png(p)}}  # noqa: E731
    txt = lambda t: {"type": "text", "text": t}  # noqa: E731
    # parts are flattened in reverse (generation.rs:251): [text, image] -> "<image>\n text"
    reqs = [
        body([txt("Free OCR."), img(pages[0])]),
        body([txt("Convert the document to markdown."), img(pages[1])]),
        body([txt("Compare."), img(pages[0]), img(pages[2])]),          # two <image> slots
        body([txt("Free OCR."), img(pages[2])], stream=True),
    ]
    results = [None] * len(reqs)

    def call(i):
        results[i] = client.post("/v1/chat/completions", json=reqs[i])

    th = [threading.Thread(target=call, args=(i,)) for i in range(len(reqs))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert all(r.status_code == 200 for r in results), [r.text for r in results]
    assert max(batcher.batches) >= 2  # concurrent requests really shared a lock-step batch
    texts = [r.json()["choices"][0]["message"]["content"] for r in results[:3]]
    deltas = [json.loads(l[6:]) for l in results[3].text.splitlines() if l.startswith("data: {")]
    streamed = "".join(c["choices"][0]["delta"].get("content", "") for c in deltas)
    assert results[3].text.strip().endswith("data: [DONE]")

    # a seeded sampling request: reproducible, and different from the greedy answer
    samp = body([txt("Free OCR."), img(pages[0])], do_sample=True, temperature=0.9, top_k=20, seed=7)
    s1 = client.post("/v1/chat/completions", json=samp).json()["choices"][0]["message"]["content"]
    s2 = client.post("/v1/chat/completions", json=samp).json()["choices"][0]["message"]["content"]
    assert s1 == s2 and s1 != texts[0]
    assert client.post("/v1/chat/completions", json=body([txt("x"), img(pages[0])], no_repeat_ngram_size=-3)).status_code == 400
    batcher.close()

    # the same requests straight through the engine, prompts / image order derived with the server's own helpers
    from dsocr.report import split_prompt_on_image, tokenize_segments
    from dsocr.server import convert_messages

    direct_in = []
    for r in reqs:
        prompt, images = convert_messages(r["messages"])
        direct_in.append((images, tokenize_segments(tok, split_prompt_on_image(prompt))))
    assert len(direct_in[2][0]) == 2 and len(direct_in[2][1]) == 3
    direct = eng.decode_requests(direct_in, VisionSettings(1024, 640, True), 2047, DecodeParameters(12, eos_token_id=1))
    eng.close()
    for i in range(3):
        assert texts[i] == tok.decode(direct[i].generated_tokens), i
        assert results[i].json()["usage"]["prompt_tokens"] == direct[i].prompt_tokens
    assert streamed == tok.decode(direct[3].generated_tokens)
