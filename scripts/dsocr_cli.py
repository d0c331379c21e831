#!/usr/bin/env python
"""CLI-equivalent driver: the flags the reference's `deepseek-ocr-cli` takes from its benchsuite
(benchsuite/models/base.py:231-296: --model --image --device --dtype --max-new-tokens [--bench --bench-output] --output-json
--prompt) plus the model / inference flags of crates/config/src/args.rs:8-86, running on the B200 engine and writing the
same `--output-json` (crates/cli/src/debug.rs:100-157) and `--bench-output` (crates/cli/src/bench.rs:138-249) files, so
the reference's strict token gate and perf tables can consume the run unchanged.

What stays on the host exactly as in the reference: image decode (PIL instead of the `image` crate), the tokenizer
(`tokenizers` JSON file), `normalize_text`.  --prompt is rendered through --template (dsocr/conversation.py == crates/core/src/conversation; the default `plain`
leaves the benchsuite's already rendered prompt unchanged apart from trimming).
Offline (no tokenizer file): --prompt-ids '[[ids before <image>], [ids after]]' replaces --prompt / --tokenizer."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))

from dsocr import report  # noqa: E402
from dsocr.conversation import render_prompt  # noqa: E402

EOS_TEXT = "<｜end▁of▁sentence｜>"


def normalize_text(s: str) -> str:
    """crates/core/src/inference.rs:228-233."""
    return s.replace("\r\n", "\n").replace(EOS_TEXT, "").strip()


def parse_bool(v: str) -> bool:
    if v.lower() in ("1", "true", "yes", "on"):
        return True
    if v.lower() in ("0", "false", "no", "off"):
        return False
    raise argparse.ArgumentTypeError(f"expected a boolean, got `{v}`")


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="dsocr-cli", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model", default="deepseek-ocr", help="model id (only `deepseek-ocr` is served by this engine)")
    ap.add_argument("--model-config", help="config.json of the checkpoint")
    ap.add_argument("--tokenizer", help="tokenizer.json")
    ap.add_argument("--weights", help="model safetensors file")
    ap.add_argument("--snapshot", help="optional DSQ snapshot (.dsq)")
    ap.add_argument("--device", default="cuda", help="cuda | cuda:N (there is no CPU path)")
    ap.add_argument("--dtype", default="bf16", choices=["f16", "bf16"])
    ap.add_argument("--template", default="plain")
    ap.add_argument("--base-size", type=int, default=1024)
    ap.add_argument("--image-size", type=int, default=640)
    ap.add_argument("--crop-mode", type=parse_bool, default=True)
    ap.add_argument("--max-new-tokens", type=int, default=512)
    ap.add_argument("--no-cache", action="store_true")
    ap.add_argument("--do-sample", type=parse_bool, default=False)
    ap.add_argument("--temperature", type=float)
    ap.add_argument("--top-p", type=float)
    ap.add_argument("--top-k", type=int)
    ap.add_argument("--repetition-penalty", type=float, default=1.0)
    ap.add_argument("--no-repeat-ngram-size", type=int, default=20)
    ap.add_argument("--seed", type=int)
    ap.add_argument("--prompt")
    ap.add_argument("--prompt-file")
    ap.add_argument("--prompt-ids", help="JSON [[ids...], [ids...]]: token ids of the text before / after <image> (offline use)")
    ap.add_argument("--image-token-id", type=int, help="id of <image> when no tokenizer file is given")
    ap.add_argument("--eos-token-id", type=int, default=1, help="config.json eos_token_id (1 for the released checkpoint)")
    ap.add_argument("--image", dest="images", action="append", default=[])
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--bench-output")
    ap.add_argument("--output-json")
    ap.add_argument("-q", "--quiet", action="store_true")
    return ap


def device_ordinal(device: str) -> int:
    d = device.lower()
    if d in ("cuda", "gpu"):
        return 0
    if d.startswith("cuda:"):
        return int(d.split(":", 1)[1])
    raise SystemExit(f"device `{device}` is not available: this engine runs on CUDA (sm_100a) only, there is no CPU or Metal path")


def resolve_prompt(args, tokenizer):
    """-> (user prompt, rendered prompt, [ids before image, ids after image], image token id)."""
    if args.prompt_ids:
        segs = json.loads(args.prompt_ids)
        if not (isinstance(segs, list) and len(segs) == len(args.images) + 1):
            raise SystemExit("prompt formatting failed: prompt/image embedding mismatch: --prompt-ids needs one id list more "
                             "than there are --image arguments (the text around every <image> slot)")
        if args.image_token_id is None:
            raise SystemExit("--prompt-ids needs --image-token-id")
        text = args.prompt or ""
        return text, text, [list(map(int, x)) for x in segs], args.image_token_id
    if args.prompt_file:
        text = Path(args.prompt_file).read_text()
    elif args.prompt is not None:
        text = args.prompt
    else:
        raise SystemExit("one of --prompt, --prompt-file or --prompt-ids is required")
    if tokenizer is None:
        raise SystemExit("--tokenizer is required to encode --prompt")
    user = text
    try:
        text = render_prompt(args.template, "", text)  # crates/cli/src/app.rs:135
    except ValueError as ex:
        raise SystemExit(f"prompt formatting failed: {ex}")
    pieces = report.split_prompt_on_image(text)
    n_images = len(args.images)
    if len(pieces) - 1 != n_images:
        # the wording the reference raises and its server maps to HTTP 400 (model/mod.rs:2550-2555)
        raise SystemExit(f"prompt formatting failed: prompt/image embedding mismatch: prompt has {len(pieces) - 1} <image> "
                         f"placeholders but {n_images} images were supplied")
    image_id = args.image_token_id if args.image_token_id is not None else tokenizer.token_to_id("<image>")
    if image_id is None:
        raise SystemExit("tokenizer has no <image> token")
    return user, text, report.tokenize_segments(tokenizer, pieces), int(image_id)


def run(args) -> int:
    import numpy as np
    from PIL import Image

    from dsocr.engine import DecodeParameters, VisionSettings, load_model

    if args.model != "deepseek-ocr":
        raise SystemExit(f"model `{args.model}` is not served by this engine (deepseek-ocr only)")
    if not args.model_config or not args.weights:
        raise SystemExit("--model-config and --weights are required")
    tokenizer = None
    if args.tokenizer:
        from tokenizers import Tokenizer
        tokenizer = Tokenizer.from_file(args.tokenizer)
    rec = report.BenchRecorder()
    t0 = time.perf_counter()
    user_prompt, rendered, segs, image_id = resolve_prompt(args, tokenizer)
    rec.record(report.STAGE_PROMPT, time.perf_counter() - t0)
    pages = [np.asarray(Image.open(p).convert("RGB")) for p in args.images]  # DynamicImage::to_rgb8 drops alpha

    t0 = time.perf_counter()
    eng = load_model(args.model_config, args.weights, args.snapshot, device_ordinal(args.device), args.dtype)
    rec.record(report.STAGE_LOAD, time.perf_counter() - t0)
    params = DecodeParameters(max_new_tokens=args.max_new_tokens, do_sample=args.do_sample, temperature=args.temperature or 0.0,
                              top_p=args.top_p, top_k=args.top_k, seed=args.seed, repetition_penalty=args.repetition_penalty,
                              no_repeat_ngram_size=args.no_repeat_ngram_size or None, eos_token_id=args.eos_token_id,
                              use_cache=not args.no_cache)
    vs = VisionSettings(args.base_size, args.image_size, args.crop_mode)
    out = eng.decode_requests([(pages, segs)], vs, image_id, params)[0]
    report.record_engine_timings(rec, eng.timings(), out.prompt_tokens, out.response_tokens)
    eng.close()

    decoded = tokenizer.decode(out.generated_tokens, skip_special_tokens=False) if tokenizer is not None else ""
    normalized = normalize_text(decoded)
    if not args.quiet:
        print(normalized if tokenizer is not None else json.dumps(out.generated_tokens))
    if args.output_json:
        report.write_output_json(args.output_json, report.CliOutput(
            model_id=args.model, weights=str(args.weights), tokenizer=str(args.tokenizer or ""), device=args.device, dtype=args.dtype,
            template=args.template, base_size=args.base_size, image_size=args.image_size, crop_mode=bool(args.crop_mode),
            max_new_tokens=args.max_new_tokens, repetition_penalty=args.repetition_penalty,
            no_repeat_ngram_size=args.no_repeat_ngram_size or None, use_cache=not args.no_cache, prompt=user_prompt,
            rendered_prompt=rendered, image_paths=[str(p) for p in args.images], prompt_tokens=out.prompt_tokens,
            generated_len=out.response_tokens, tokens=out.generated_tokens, decoded=decoded, normalized=normalized))
    if args.bench or args.bench_output:
        if args.bench_output:
            rec.write(args.bench_output)
        elif not args.quiet:
            print(json.dumps(rec.to_json()["stage_totals"]), file=sys.stderr)
    return 0


def main(argv=None) -> int:
    return run(build_parser().parse_args(argv))


if __name__ == "__main__":
    sys.exit(main())
