"""OpenAI-compatible `/v1/chat/completions` and `/v1/responses` (+ `/v1/models`, `/v1/health`) on top of the batcher: the routes of
crates/server/src/routes.rs:33-222 with the message handling of generation.rs:169-300 (latest user message plus the
system messages before it; content parts are flattened in reverse order, images become `<image>` placeholders; only
`data:` URLs carry images here - this box has no egress for http(s) ones) and the SSE chunks of stream.rs:160-370
(`chat.completion.chunk`: a role chunk, UTF-8-safe content deltas from DeltaTracker, a `finish_reason: "stop"` chunk with
usage, `[DONE]`).  Unlike the reference, concurrent requests do not queue on an engine mutex: they are batched."""
from __future__ import annotations

import asyncio
import base64
import io
import json
import time
import uuid
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
from fastapi import FastAPI, Request
from fastapi.responses import JSONResponse, StreamingResponse

from .batcher import PageBatcher, PageRequest
from .report import split_prompt_on_image, tokenize_segments
from .streaming import DeltaTracker

MISSING_IMAGE = ("⚠️ **Image Required**\n\n- This OCR backend expects at least one `<image>` placeholder or attached image.\n"
                 "- Please include `input_image` / `image_url`, or add `<image>` inside the prompt.\n\n---\n\n"
                 "⚠️ **需要图像输入**\n\n- 当前 OCR 模型需要至少一个 `<image>` 占位符或实际图片。\n"
                 "- 请在请求中附带 `input_image`/`image_url`，或在 prompt 中插入 `<image>`。")
EOS_TEXT = "<｜end▁of▁sentence｜>"


class BadRequest(Exception):
    pass


def load_image(url: str) -> np.ndarray:
    from PIL import Image

    if url.startswith("data:"):
        meta, sep, payload = url[5:].partition(",")
        if not sep:
            raise BadRequest("invalid data URL")
        if not meta.endswith(";base64"):
            raise BadRequest("data URLs must specify base64 encoding")
        try:
            raw = base64.b64decode(payload, validate=True)
            return np.asarray(Image.open(io.BytesIO(raw)).convert("RGB"))
        except Exception as ex:
            raise BadRequest(f"failed to decode inline image: {ex}")
    if url.startswith("http://") or url.startswith("https://"):
        raise BadRequest("remote image URLs are not reachable from this server (no egress); send a data: URL")
    raise BadRequest("only data: URIs or http(s) image URLs are supported")


def flatten_content(content: Any) -> Tuple[str, List[np.ndarray]]:
    """generation.rs:246-268."""
    if content is None:
        return "", []
    if isinstance(content, str):
        return content.strip(), []
    buf, images = "", []
    for part in reversed(content):
        kind = part.get("type")
        if kind in ("image_url", "input_image"):
            spec = part.get("image_url")
            buf += "<image>"
            images.append(load_image(spec if isinstance(spec, str) else (spec or {}).get("url", "")))
        elif kind in ("text", "input_text"):
            if buf:
                buf += "\n"
            buf += part.get("text", "")
        else:
            raise BadRequest(f"unsupported content part type `{kind}`")
    return buf.strip(), images


def convert_messages(messages: List[Dict[str, Any]]) -> Tuple[str, List[np.ndarray]]:
    """generation.rs:181-243: only one round of prompt is kept (the model is not trained for dialogue)."""
    idx = None
    for i, m in enumerate(messages):
        if str(m.get("role", "")).lower() == "user":
            idx = i
    if idx is None:
        raise BadRequest("request must include at least one user message")
    sections, images = [], []
    for m in messages[:idx]:
        if str(m.get("role", "")).lower() != "system":
            continue
        text, imgs = flatten_content(m.get("content"))
        if text:
            sections.append(text)
        images += imgs
    text, imgs = flatten_content(messages[idx].get("content"))
    if text:
        sections.append(text)
    images += imgs
    if not sections and not images:
        raise BadRequest("user content must include text or images")
    return "\n\n".join(sections).strip(), images


def params_from_tuple(p: Tuple):
    """The hashable decode-parameter tuple a request carries through the batcher -> dsocr.engine.DecodeParameters."""
    from .engine import DecodeParameters

    budget, ngram, eos, do_sample, temperature, top_p, top_k, penalty, seed = p
    return DecodeParameters(max_new_tokens=budget, no_repeat_ngram_size=ngram or None, eos_token_id=eos, do_sample=do_sample,
                            temperature=temperature, top_p=top_p, top_k=top_k, repetition_penalty=penalty, seed=seed)


def create_app(batcher: PageBatcher, tokenizer: Any, image_token_id: int, model_id: str = "deepseek-ocr",
               vision: Tuple[int, int, bool] = (1024, 640, True), max_new_tokens: int = 512, no_repeat_ngram_size: int = 20,
               max_budget: int = 8192 - 1024):
    app = FastAPI(title="dsocr-b200")

    def decode_text(ids: List[int]) -> str:
        return tokenizer.decode(ids, skip_special_tokens=False) if ids else ""

    def final_text(ids: List[int]) -> str:
        return decode_text(ids).replace("\r\n", "\n").replace(EOS_TEXT, "").strip()  # normalize_text (inference.rs:228-233)

    @app.get("/v1/health")
    def health():
        return {"status": "ok"}

    @app.get("/v1/models")
    def models():
        return {"object": "list", "data": [{"id": model_id, "object": "model", "created": 0, "owned_by": "dsocr-b200"}]}

    def sse(obj: Dict[str, Any]) -> str:
        return f"data: {json.dumps(obj, ensure_ascii=False)}\n\n"

    class ChatShape:
        """`/v1/chat/completions` bodies and chunks (routes.rs:142-222, stream.rs StreamKind::Chat)."""
        messages_key = "messages"

        def __init__(self):
            self.id, self.created = f"chatcmpl-{uuid.uuid4()}", int(time.time())

        def budget(self, req):
            return req.get("max_tokens")

        def body(self, text, p, c):
            return {"id": self.id, "object": "chat.completion", "created": self.created, "model": model_id,
                    "choices": [{"index": 0, "message": {"role": "assistant", "content": text}, "finish_reason": "stop"}],
                    "usage": {"prompt_tokens": p, "completion_tokens": c, "total_tokens": p + c}}

        def _chunk(self, delta, finish, usage=None):
            o = {"id": self.id, "object": "chat.completion.chunk", "created": self.created, "model": model_id,
                 "choices": [{"index": 0, "delta": delta, "finish_reason": finish}]}
            if usage is not None:
                o["usage"] = usage
            return sse(o)

        def initial(self):
            return self._chunk({"role": "assistant"}, None)

        def delta(self, text):
            return self._chunk({"content": text}, None)

        def final(self, text, p, c):
            return self._chunk({}, "stop", {"prompt_tokens": p, "completion_tokens": c, "total_tokens": p + c})

        def error(self):
            return self._chunk({}, "error")

    class ResponsesShape:
        """`/v1/responses` bodies and events (routes.rs:54-140, stream.rs StreamKind::Responses)."""
        messages_key = "input"

        def __init__(self):
            self.id, self.out_id, self.created = f"resp-{uuid.uuid4()}", f"msg-{uuid.uuid4()}", int(time.time())

        def budget(self, req):
            return req.get("max_output_tokens") or req.get("max_tokens")

        def _head(self):
            return {"id": self.id, "object": "response", "created": self.created, "model": model_id}

        def _output(self, text):
            return [{"id": self.out_id, "type": "message", "role": "assistant", "content": [{"type": "output_text", "text": text}]}]

        def body(self, text, p, c):
            return {**self._head(), "output": self._output(text),
                    "usage": {"prompt_tokens": p, "completion_tokens": c, "total_tokens": p + c}}

        def initial(self):
            return sse({"type": "response.created", "response": self._head()})

        def delta(self, text):
            return sse({"type": "response.output_text.delta", "response": self._head(), "output_id": self.out_id, "output_index": 0,
                        "delta": text})

        def final(self, text, p, c):
            return sse({"type": "response.completed", "response": {**self._head(), "output": self._output(text),
                                                                   "usage": {"input_tokens": p, "output_tokens": c, "total_tokens": p + c}}})

        def error(self):
            return sse({"type": "response.error", "response": self._head()})

    def bad(message: str, code: int = 400):
        return JSONResponse({"error": {"message": message, "type": "invalid_request_error" if code == 400 else "server_error"}},
                            status_code=code)

    async def generate(request: Request, shape):
        try:
            req = await request.json()
        except Exception:
            return bad("invalid JSON body")
        if req.get("model") not in (None, "", model_id):
            return bad(f"requested model `{req.get('model')}` is not available")
        try:
            prompt, images = convert_messages(req.get(shape.messages_key) or [])
        except BadRequest as ex:
            return bad(str(ex))
        stream = bool(req.get("stream"))

        if "<image>" not in prompt:  # prompt_missing_image (routes.rs:241-247)
            if not stream:
                return shape.body(MISSING_IMAGE, 0, 0)

            async def fallback():
                yield shape.initial()
                yield shape.delta(MISSING_IMAGE)
                yield shape.final(MISSING_IMAGE, 0, 0)
                yield "data: [DONE]\n\n"
            return StreamingResponse(fallback(), media_type="text/event-stream")

        pieces = split_prompt_on_image(prompt)
        if len(pieces) - 1 != len(images):
            # the wording the reference's server maps to HTTP 400 (generation.rs:111-115, model/mod.rs:2550-2555)
            return bad(f"prompt formatting failed: prompt/image embedding mismatch: {len(pieces) - 1} slots vs "
                       f"{len(images)} embeddings")
        segs = tokenize_segments(tokenizer, pieces)
        try:
            # the budget is clamped to what the server allows (a larger one would only fail later against the position table)
            budget = max(0, min(int(shape.budget(req) or max_new_tokens), max_budget))
            ngram = int(req["no_repeat_ngram_size"]) if req.get("no_repeat_ngram_size") is not None else no_repeat_ngram_size
            if ngram < 0 or ngram > 4096:
                return bad("no_repeat_ngram_size must be between 0 and 4096")
            # DecodeParameters of the request (routes.rs / generation.rs forward them unchanged): greedy unless do_sample
            temperature = float(req.get("temperature") or 0.0)
            do_sample = bool(req.get("do_sample", False))
            top_p = float(req["top_p"]) if req.get("top_p") is not None else 1.0
            top_k = int(req["top_k"]) if req.get("top_k") is not None else None
            penalty = float(req.get("repetition_penalty") or 1.0)
            seed = int(req["seed"]) if req.get("seed") is not None else None
        except (TypeError, ValueError) as ex:
            return bad(f"invalid generation parameter: {ex}")
        params = (budget, ngram, 1, do_sample, temperature, top_p, top_k, penalty, seed)
        loop = asyncio.get_running_loop()
        events: "asyncio.Queue[Tuple[int, List[int]]]" = asyncio.Queue()

        def on_tokens(count: int, tokens: List[int]):  # called on the batcher thread
            loop.call_soon_threadsafe(events.put_nowait, (count, list(tokens)))

        pr = PageRequest(page=images[0], seg0=tuple(segs[0]), seg1=tuple(segs[1]), image_token_id=image_token_id, vision=vision,
                         params=params, on_tokens=on_tokens if stream else None,
                         more=tuple((images[i], tuple(segs[i + 1])) for i in range(1, len(images))))
        fut = asyncio.wrap_future(batcher.submit(pr))
        if not stream:
            try:
                out = await fut
            except Exception as ex:
                client_fault = "prompt formatting failed" in str(ex) or "embedding mismatch" in str(ex)
                return bad(str(ex), 400 if client_fault else 500)
            return shape.body(final_text(out.generated_tokens), out.prompt_tokens, out.response_tokens)

        async def events_stream():
            tracker = DeltaTracker()
            yield shape.initial()
            while True:
                getter = asyncio.ensure_future(events.get())
                done, _ = await asyncio.wait({getter, fut}, return_when=asyncio.FIRST_COMPLETED)
                if getter in done:
                    _, toks = getter.result()
                    delta = tracker.advance(decode_text(toks), False)
                    if delta:
                        yield shape.delta(delta)
                    continue
                getter.cancel()
                break
            try:
                out = fut.result()
            except Exception:
                yield shape.error()
                yield "data: [DONE]\n\n"
                return
            text = final_text(out.generated_tokens)
            delta = tracker.advance(text, True)  # whatever was still held back (and the normalisation of the tail)
            if delta:
                yield shape.delta(delta)
            yield shape.final(text, out.prompt_tokens, out.response_tokens)
            yield "data: [DONE]\n\n"
        return StreamingResponse(events_stream(), media_type="text/event-stream")

    @app.post("/v1/chat/completions")
    async def chat(request: Request):
        return await generate(request, ChatShape())

    @app.post("/v1/responses")
    async def responses(request: Request):
        return await generate(request, ResponsesShape())

    return app
