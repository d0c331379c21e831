"""Host-side run reports and the token gate (dsocr/report.py, dsocr/gate.py): same JSON schemas as the reference CLI's
`--output-json` / `--bench-output` (crates/cli/src/debug.rs:100-157, bench.rs:138-249) and the same comparison rule as
the reference's benchsuite.  Golden expectations come from the reference's own Python (tests/golden/make_gate_golden.py);
when /root/reference is present the emitted files are also parsed with the reference's own schema classes."""
import json
import os
import sys

import pytest

from dsocr import gate, report

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "gate_cases.json")
HAVE_REF = os.path.isdir("/root/reference/benchsuite")


def test_strict_compare_matches_reference_outputs():
    cases = json.load(open(GOLDEN))
    assert len(cases) >= 9
    for c in cases:
        assert gate.strict_compare(c["python"], c["rust"]) == c["expected"], c


def test_token_agreement_uses_the_gate_trim():
    assert gate.token_agreement([5, 6, 7, 1], [5, 6, 7]) == 1.0
    assert gate.token_agreement([5, 6, 7, 8], [5, 6, 9, 8]) == 0.75
    assert gate.token_agreement([], [1]) == 1.0


def _sample_output():
    return report.CliOutput(
        model_id="deepseek-ocr", weights="/w/model.safetensors", tokenizer="/w/tokenizer.json", device="cuda:0", dtype="bf16",
        template="plain", base_size=1024, image_size=640, crop_mode=True, max_new_tokens=512, repetition_penalty=1.0,
        no_repeat_ngram_size=20, use_cache=True, prompt="<image>\nFree OCR.", rendered_prompt="<image>\nFree OCR.",
        image_paths=["page.png"], prompt_tokens=913, generated_len=3, tokens=[11, 12, 13], decoded="abc", normalized="abc")


def test_output_json_has_the_reference_fields_in_order(tmp_path):
    p = tmp_path / "sub" / "rust_output.json"
    report.write_output_json(str(p), _sample_output())
    d = json.loads(p.read_text())
    assert list(d) == ["schema_version", "model_id", "weights", "tokenizer", "device", "dtype", "template", "base_size", "image_size",
                       "crop_mode", "max_new_tokens", "repetition_penalty", "no_repeat_ngram_size", "use_cache", "prompt",
                       "rendered_prompt", "image_paths", "prompt_tokens", "generated_len", "tokens", "decoded", "normalized"]
    assert d["schema_version"] == 1 and d["tokens"] == [11, 12, 13] and d["no_repeat_ngram_size"] == 20


def test_bench_report_totals(tmp_path):
    rec = report.BenchRecorder()
    rec.record(report.STAGE_LOAD, 1.5)
    rec.record_ms(report.STAGE_PREFILL, 40.0, prompt_tokens=913)
    rec.record_ms(report.STAGE_ITERATIVE, 300.0, generated_tokens=512)
    rec.record_ms(report.STAGE_ITERATIVE, 100.0)
    p = tmp_path / "bench_raw.json"
    rec.write(str(p))
    d = json.loads(p.read_text())
    assert [e["stage"] for e in d["events"]] == ["model.load", "decode.prefill", "decode.iterative", "decode.iterative"]
    assert d["events"][1]["fields"] == [{"key": "prompt_tokens", "value": 913}]
    assert d["events"][0]["duration_ns"] == "1500000000" and d["events"][0]["duration_ms"] == 1500.0
    it = {s["stage"]: s for s in d["stage_totals"]}["decode.iterative"]
    assert it["count"] == 2 and it["total_ms"] == 400.0 and it["avg_ms"] == 200.0 and it["min_ms"] == 100.0 and it["max_ms"] == 300.0
    assert it["total_ns"] == "400000000"


def test_engine_timings_become_reference_stage_events():
    rec = report.BenchRecorder()
    report.record_engine_timings(rec, {"vision.prepare_inputs": 4.5, "vision.compute_embeddings": 125.0, "decode.prefill": 43.0,
                                       "decode.iterative": 880.0, "decode.generate": 923.0}, prompt_tokens=285, generated=512)
    assert [e.stage for e in rec.events] == ["vision.prepare_inputs", "vision.compute_embeddings", "decode.prefill",
                                             "decode.iterative", "decode.generate"]
    assert rec.events[2].fields == {"prompt_tokens": 285}


def test_prompt_split_and_tokenisation():
    class Tok:  # test double with the `tokenizers.Tokenizer` call shape
        def encode(self, text, add_special_tokens=False):
            assert add_special_tokens is False
            return type("Enc", (), {"ids": [len(w) for w in text.split()]})()

    assert report.split_prompt_on_image("<image>\nFree OCR.") == ["", "\nFree OCR."]
    assert report.split_prompt_on_image("a <image> b <image>") == ["a ", " b ", ""]
    assert report.tokenize_segments(Tok(), ["", "\nFree OCR."]) == [[], [4, 4]]


@pytest.mark.skipif(not HAVE_REF, reason="reference benchsuite not present on this box")
def test_emitted_files_parse_with_the_reference_schema_classes(tmp_path):
    sys.path.insert(0, "/root/reference")
    from benchsuite.schemas import RustDecodeOutput, StageTotals

    report.write_output_json(str(tmp_path / "o.json"), _sample_output())
    out = RustDecodeOutput.from_payload(json.loads((tmp_path / "o.json").read_text()), token_field="tokens")
    assert out.tokens == [11, 12, 13] and out.prompt_tokens == 913 and out.generated_len == 3 and out.rendered_prompt == "<image>\nFree OCR."
    rec = report.BenchRecorder()
    rec.record(report.STAGE_LOAD, 2.0)
    rec.record_ms(report.STAGE_PREFILL, 43.0)
    rec.record_ms(report.STAGE_ITERATIVE, 880.0)
    rec.record_ms(report.STAGE_GENERATE, 923.0)
    rec.write(str(tmp_path / "b.json"))
    st = StageTotals.from_payload(json.loads((tmp_path / "b.json").read_text()))
    assert st.stage_ms("model.load") == 2000.0 and st.stage_ms("decode.iterative") == 880.0 and st.stage_ms("missing") == 0.0
