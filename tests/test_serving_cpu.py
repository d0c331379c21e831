"""Host-side serving pieces without a GPU: the UTF-8-safe delta tracker (crates/core/src/streaming.rs)
and the request batcher that replaces the reference's one-request-at-a-time engine mutex."""
import threading
import time

import numpy as np

from dsocr.batcher import PageBatcher, PageRequest
from dsocr.streaming import DeltaTracker, extract_delta


def test_extract_delta():
    assert extract_delta("abc", "abcdef") == "def"
    assert extract_delta("abc", "abXdef") == "Xdef"        # diverging history: resend from the first difference
    assert extract_delta("", "x") == "x" and extract_delta("same", "same") == ""
    assert extract_delta("日本", "日本語") == "語"


def test_delta_tracker_holds_back_incomplete_utf8():
    t = DeltaTracker()
    assert t.advance("Hel", False) == "Hel"
    assert t.advance("Hello �", False) == "lo "       # trailing replacement char: emit only what is complete
    assert t.snapshot() == "Hello "
    assert t.advance("Hello �", False) == ""          # nothing new that is complete
    assert t.advance("Hello 世", False) == "世"
    assert t.advance("Hello 世�", True) == "�"   # final call: everything goes through
    t.reset()
    assert t.snapshot() == "" and t.advance("�", False) == ""


def _req(i, key="a", on_tokens=None):
    # requests batch together when vision settings / decode parameters agree (prompts may differ per request)
    return PageRequest(page=np.zeros((2, 2, 3), np.uint8), seg0=(1,), seg1=(2, 3) if key == "a" else (9,), image_token_id=7,
                       vision=(1024, 640, True), params=(64, 20, 1) if key == "a" else (32, 20, 1), on_tokens=on_tokens)


def test_batcher_groups_concurrent_compatible_requests():
    seen = []

    def run(batch):
        seen.append([r.params for r in batch])
        time.sleep(0.01)
        for r in batch:
            if r.on_tokens:
                r.on_tokens(2, [5, 6])
        return [("ok", r.seg1) for r in batch]

    b = PageBatcher(run, max_batch=4, max_wait_ms=50)
    streamed = []
    futs = [b.submit(_req(i, "a", on_tokens=(lambda c, t: streamed.append((c, t))) if i == 0 else None)) for i in range(6)]
    futs += [b.submit(_req(9, "b"))]
    res = [f.result(timeout=5) for f in futs]
    b.close()
    assert [r[0] for r in res] == ["ok"] * 7 and res[-1][1] == (9,)
    assert sorted(b.batches, reverse=True)[:2] == [4, 2] and sum(b.batches) == 7    # 6 compatible -> 4 + 2, the odd one alone
    assert all(len(set(s)) == 1 for s in seen)                                      # never mixes incompatible requests
    assert streamed == [(2, [5, 6])]


def test_batcher_propagates_engine_errors_and_keeps_serving():
    calls = {"n": 0}

    def run(batch):
        calls["n"] += 1
        if calls["n"] == 1:
            raise RuntimeError("prompt formatting failed: prompt/image embedding mismatch")
        return [len(batch)] * len(batch)

    b = PageBatcher(run, max_batch=8, max_wait_ms=1)
    f1 = b.submit(_req(0))
    try:
        f1.result(timeout=5)
        assert False
    except RuntimeError as ex:
        assert "embedding mismatch" in str(ex)
    assert b.submit(_req(1)).result(timeout=5) == 1
    b.close()


def test_batcher_many_threads():
    def run(batch):
        time.sleep(0.002)
        return [r.image_token_id for r in batch]

    b = PageBatcher(run, max_batch=16, max_wait_ms=10)
    out = []
    lock = threading.Lock()

    def client():
        r = b.submit(_req(0)).result(timeout=10)
        with lock:
            out.append(r)

    th = [threading.Thread(target=client) for _ in range(40)]
    [t.start() for t in th]
    [t.join() for t in th]
    b.close()
    assert out == [7] * 40 and sum(b.batches) == 40 and max(b.batches) > 1 and max(b.batches) <= 16


# ------------------------------------------------------------------------------------------------ HTTP layer
class _Tok:
    def encode(self, text, add_special_tokens=False):
        return type("Enc", (), {"ids": [100 + len(w) for w in text.split()]})()

    def decode(self, ids, skip_special_tokens=False):
        return "".join(chr(0x4E00 + i) if i < 100 else "�" for i in ids)  # ids >= 100 decode to an incomplete character


def _png_data_url():
    import base64
    import io

    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(np.full((8, 6, 3), 200, np.uint8)).save(buf, format="PNG")
    return "data:image/png;base64," + base64.b64encode(buf.getvalue()).decode()


def _client(run):
    from fastapi.testclient import TestClient

    from dsocr.server import create_app

    b = PageBatcher(run, max_batch=8, max_wait_ms=2)
    return TestClient(create_app(b, _Tok(), image_token_id=777)), b


def test_chat_completion_roundtrip_and_message_flattening():
    from dsocr.engine import DecodeOutcome

    seen = {}

    def run(batch):
        r = batch[0]
        seen.update(seg0=r.seg0, seg1=r.seg1, shape=r.page.shape, params=r.params, image_id=r.image_token_id)
        return [DecodeOutcome(281, 3, [1, 2, 3]) for _ in batch]

    client, b = _client(run)
    body = {"model": "deepseek-ocr", "max_tokens": 32, "messages": [
        {"role": "system", "content": "system text before the user turn is kept"},
        {"role": "assistant", "content": "dropped"},
        {"role": "user", "content": [{"type": "text", "text": "Free OCR."}, {"type": "image_url", "image_url": {"url": _png_data_url()}}]}]}
    r = client.post("/v1/chat/completions", json=body)
    assert r.status_code == 200, r.text
    d = r.json()
    assert d["object"] == "chat.completion" and d["choices"][0]["message"] == {"role": "assistant", "content": "".join(chr(0x4E00 + i) for i in (1, 2, 3))}
    assert d["usage"] == {"prompt_tokens": 281, "completion_tokens": 3, "total_tokens": 284}
    # parts are flattened in reverse (generation.rs:251): "<image>\nFree OCR." after the system section
    assert seen["shape"] == (8, 6, 3) and seen["params"][:3] == (32, 20, 1) and seen["params"][3] is False and seen["image_id"] == 777
    assert seen["seg1"] == (104, 104) and len(seen["seg0"]) == 8  # 8 words of the system section, then <image>
    assert client.get("/v1/models").json()["data"][0]["id"] == "deepseek-ocr" and client.get("/v1/health").json() == {"status": "ok"}
    b.close()


def test_chat_errors_and_missing_image_fallback():
    client, b = _client(lambda batch: [None] * len(batch))
    r = client.post("/v1/chat/completions", json={"model": "deepseek-ocr", "messages": [{"role": "assistant", "content": "x"}]})
    assert r.status_code == 400 and "at least one user message" in r.json()["error"]["message"]
    r = client.post("/v1/chat/completions", json={"model": "other", "messages": [{"role": "user", "content": "x"}]})
    assert r.status_code == 400 and "not available" in r.json()["error"]["message"]
    r = client.post("/v1/chat/completions", json={"model": "deepseek-ocr", "messages": [{"role": "user", "content": "no image here"}]})
    assert r.status_code == 200 and "Image Required" in r.json()["choices"][0]["message"]["content"]
    r = client.post("/v1/chat/completions", json={"model": "deepseek-ocr", "messages": [{"role": "user", "content": [
        {"type": "image_url", "image_url": "http://example.com/a.png"}]}]})
    assert r.status_code == 400 and "no egress" in r.json()["error"]["message"]
    b.close()


def test_chat_streaming_emits_role_deltas_stop_and_done():
    import json as _json

    from dsocr.engine import DecodeOutcome

    def run(batch):
        for r in batch:
            r.on_tokens(1, [1])
            time.sleep(0.02)
            r.on_tokens(2, [1, 150])     # second token is half a character: nothing new may be sent
            time.sleep(0.02)
            r.on_tokens(3, [1, 2, 3])
            time.sleep(0.02)
        return [DecodeOutcome(281, 3, [1, 2, 3]) for _ in batch]

    client, b = _client(run)
    body = {"model": "deepseek-ocr", "stream": True, "messages": [{"role": "user", "content": [
        {"type": "text", "text": "Free OCR."}, {"type": "image_url", "image_url": _png_data_url()}]}]}
    with client.stream("POST", "/v1/chat/completions", json=body) as r:
        assert r.status_code == 200 and r.headers["content-type"].startswith("text/event-stream")
        lines = [l for l in r.iter_lines() if l.startswith("data: ")]
    assert lines[-1] == "data: [DONE]"
    chunks = [_json.loads(l[6:]) for l in lines[:-1]]
    assert all(c["object"] == "chat.completion.chunk" for c in chunks)
    assert chunks[0]["choices"][0]["delta"] == {"role": "assistant"}
    text = "".join(c["choices"][0]["delta"].get("content", "") for c in chunks)
    assert text == "".join(chr(0x4E00 + i) for i in (1, 2, 3)) and "�" not in text
    assert chunks[-1]["choices"][0]["finish_reason"] == "stop" and chunks[-1]["usage"]["total_tokens"] == 284
    b.close()


def test_responses_route_json_and_stream():
    import json as _json

    from dsocr.engine import DecodeOutcome

    def run(batch):
        for r in batch:
            if r.on_tokens:
                r.on_tokens(2, [4, 5])
                time.sleep(0.02)
        return [DecodeOutcome(281, 2, [4, 5]) for _ in batch]

    client, b = _client(run)
    body = {"model": "deepseek-ocr", "max_output_tokens": 16, "input": [{"role": "user", "content": [
        {"type": "input_text", "text": "Free OCR."}, {"type": "input_image", "image_url": _png_data_url()}]}]}
    d = client.post("/v1/responses", json=body).json()
    text = "".join(chr(0x4E00 + i) for i in (4, 5))
    assert d["object"] == "response" and d["output"][0]["content"] == [{"type": "output_text", "text": text}]
    assert d["usage"] == {"prompt_tokens": 281, "completion_tokens": 2, "total_tokens": 283}
    with client.stream("POST", "/v1/responses", json={**body, "stream": True}) as r:
        lines = [l for l in r.iter_lines() if l.startswith("data: ")]
    assert lines[-1] == "data: [DONE]"
    ev = [_json.loads(l[6:]) for l in lines[:-1]]
    assert ev[0]["type"] == "response.created" and ev[-1]["type"] == "response.completed"
    assert "".join(e["delta"] for e in ev if e["type"] == "response.output_text.delta") == text
    assert ev[-1]["response"]["usage"] == {"input_tokens": 281, "output_tokens": 2, "total_tokens": 283}
    assert ev[-1]["response"]["output"][0]["content"][0]["text"] == text
    b.close()
