"""scripts/dsocr_cli.py (the CLI-equivalent driver) without a GPU: flag surface the reference's benchsuite
uses, prompt handling, error wording, and - with the engine replaced by a stub - the files it writes, parsed back with the
reference's own schema classes when /root/reference is present."""
import importlib.util
import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("dsocr_cli", ROOT / "scripts" / "dsocr_cli.py")
cli = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cli)


class Tok:
    def encode(self, text, add_special_tokens=False):
        return type("Enc", (), {"ids": [100 + len(w) for w in text.split()]})()

    def token_to_id(self, t):
        return 777 if t == "<image>" else None

    def decode(self, ids, skip_special_tokens=False):
        return " ".join(f"t{i}" for i in ids) + cli.EOS_TEXT + "\r\n"


def test_benchsuite_command_line_is_accepted():
    """benchsuite/models/base.py:262-283 (run_rust_bench)."""
    a = cli.build_parser().parse_args(["--model", "deepseek-ocr", "--image", "page.png", "--device", "cuda", "--dtype", "bf16",
                                       "--max-new-tokens", "64", "--bench", "--bench-output", "b.json", "--output-json", "o.json",
                                       "--prompt", "<image>\nFree OCR."])
    assert a.images == ["page.png"] and a.bench and a.max_new_tokens == 64 and a.crop_mode is True and a.no_repeat_ngram_size == 20


def test_normalize_text_and_device():
    assert cli.normalize_text(" a\r\nb" + cli.EOS_TEXT + " ") == "a\nb"
    assert cli.device_ordinal("cuda") == 0 and cli.device_ordinal("cuda:3") == 3
    with pytest.raises(SystemExit, match="no CPU"):
        cli.device_ordinal("cpu")


def test_prompt_resolution():
    p = cli.build_parser()
    a = p.parse_args(["--image", "x.png", "--prompt", "<image>\nFree OCR."])
    user, rendered, segs, image_id = cli.resolve_prompt(a, Tok())
    assert rendered == "<image>\nFree OCR." and segs == [[], [104, 104]] and image_id == 777
    with pytest.raises(SystemExit, match="prompt/image embedding mismatch"):
        cli.resolve_prompt(p.parse_args(["--image", "x.png", "--prompt", "no placeholder"]), Tok())
    a = p.parse_args(["--image", "x.png", "--prompt-ids", "[[5],[6,7]]", "--image-token-id", "9"])
    assert cli.resolve_prompt(a, None)[2:] == ([[5], [6, 7]], 9)


def test_run_writes_reference_schema_files(tmp_path, monkeypatch):
    from PIL import Image

    import dsocr.engine as E

    calls = {}

    class StubEngine:
        def decode_requests(self, requests, vs, image_id, params):
            (pages, (seg0, seg1)), = requests
            calls.update(shape=pages[0].shape, seg0=list(seg0), seg1=list(seg1), image_id=image_id, max_new=params.max_new_tokens,
                         ngram=params.no_repeat_ngram_size, vs=(vs.base_size, vs.image_size, vs.crop_mode))
            return [E.DecodeOutcome(913, 3, [11, 12, 13])]

        def timings(self):
            return {"vision.prepare_inputs": 4.0, "vision.compute_embeddings": 120.0, "decode.prefill": 40.0,
                    "decode.iterative": 2.0, "decode.generate": 42.0}

        def close(self):
            pass

    monkeypatch.setattr(E, "load_model", lambda *a, **k: StubEngine())
    import tokenizers
    monkeypatch.setattr(tokenizers.Tokenizer, "from_file", staticmethod(lambda path: Tok()))
    img = tmp_path / "page.png"
    Image.fromarray(np.full((50, 40, 4), 200, np.uint8), "RGBA").save(img)
    out_json, bench_json = tmp_path / "o" / "rust_output.json", tmp_path / "o" / "bench_raw.json"
    rc = cli.main(["--model", "deepseek-ocr", "--image", str(img), "--device", "cuda", "--dtype", "bf16", "--max-new-tokens", "64",
                   "--bench", "--bench-output", str(bench_json), "--output-json", str(out_json), "--prompt", "<image>\nFree OCR.",
                   "--model-config", "c.json", "--weights", "w.safetensors", "--tokenizer", "tok.json", "--quiet"])
    assert rc == 0
    assert calls == {"shape": (50, 40, 3), "seg0": [], "seg1": [104, 104], "image_id": 777, "max_new": 64, "ngram": 20,
                     "vs": (1024, 640, True)}
    o = json.loads(out_json.read_text())
    assert o["schema_version"] == 1 and o["tokens"] == [11, 12, 13] and o["prompt_tokens"] == 913 and o["generated_len"] == 3
    assert o["rendered_prompt"] == "<image>\nFree OCR." and o["normalized"] == "t11 t12 t13" and o["use_cache"] is True
    b = json.loads(bench_json.read_text())
    stages = {s["stage"]: s["total_ms"] for s in b["stage_totals"]}
    assert stages["decode.prefill"] == 40.0 and stages["decode.generate"] == 42.0 and "model.load" in stages and "prompt.render" in stages
    if os.path.isdir("/root/reference/benchsuite"):
        sys.path.insert(0, "/root/reference")
        from benchsuite.schemas import RustDecodeOutput, StageTotals
        r = RustDecodeOutput.from_payload(o, token_field="tokens")
        assert r.tokens == [11, 12, 13] and r.generated_len == 3
        assert StageTotals.from_payload(b).stage_ms("decode.iterative") == 2.0


def test_conversation_templates():
    """crates/core/tests/conversation_templates.rs:3-18 + render_prompt (inference.rs:212-225)."""
    from dsocr.conversation import get_conv_template, render_prompt

    conv = get_conv_template("deepseek")
    for i, m in enumerate(["Hello!", "Hi! This is Tony.", "Who are you?", "I am a helpful assistant.", "How are you?", None]):
        conv.append_message(conv.roles[i % 2], m)
    prompt = conv.get_prompt()
    assert "Hello!" in prompt and "<｜end▁of▁sentence｜>" in prompt
    assert prompt.startswith("<|User|>: Hello!\n\n<|Assistant|>: Hi! This is Tony.<｜end▁of▁sentence｜>") and prompt.endswith("<|Assistant|>:")
    assert get_conv_template("deepseek").messages == []             # registry hands out copies
    assert render_prompt("plain", "", "<image>\nFree OCR. ") == "<image>\nFree OCR."
    assert render_prompt("deepseek", "", "<image>\nFree OCR.") == "User: <image>\nFree OCR.\n\nAssistant:"
    assert render_prompt("alignment", "", "anything") == "<image>\n"
    assert get_conv_template("nope") is None
    with pytest.raises(ValueError, match="unknown conversation template"):
        render_prompt("nope", "", "x")


def test_cli_renders_prompt_through_template():
    p = cli.build_parser()
    a = p.parse_args(["--image", "x.png", "--prompt", "  <image>\nFree OCR.  ", "--template", "plain"])
    user, rendered, segs, _ = cli.resolve_prompt(a, Tok())
    assert user == "  <image>\nFree OCR.  " and rendered == "<image>\nFree OCR." and segs == [[], [104, 104]]
    with pytest.raises(SystemExit, match="unknown conversation template"):
        cli.resolve_prompt(p.parse_args(["--image", "x.png", "--prompt", "<image>", "--template", "nope"]), Tok())
