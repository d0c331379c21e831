#!/usr/bin/env python
"""`deepseek-ocr-server` equivalent (crates/server): OpenAI-compatible /v1/chat/completions, /v1/models, /v1/health on the
B200 engine, with concurrent requests batched into one lock-step decode instead of queueing on an engine mutex.
  python scripts/dsocr_server.py --model-config config.json --weights model.safetensors --tokenizer tokenizer.json \\
      [--snapshot model.q4k.dsq] [--device cuda:0] [--dtype bf16] [--host 0.0.0.0] [--port 8000] [--max-batch 64] [--max-wait-ms 5]"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model-config", required=True)
    ap.add_argument("--weights", required=True)
    ap.add_argument("--tokenizer", required=True)
    ap.add_argument("--snapshot")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--dtype", default="bf16", choices=["f16", "bf16"])
    ap.add_argument("--base-size", type=int, default=1024)
    ap.add_argument("--image-size", type=int, default=640)
    ap.add_argument("--crop-mode", default="true")
    ap.add_argument("--max-new-tokens", type=int, default=512)
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--max-batch", type=int, default=64)
    ap.add_argument("--max-wait-ms", type=float, default=5.0)
    a = ap.parse_args(argv)

    import uvicorn
    from tokenizers import Tokenizer

    from dsocr.batcher import PageBatcher, engine_runner
    from dsocr.engine import DecodeParameters, VisionSettings, load_model
    from dsocr.server import create_app, params_from_tuple

    tok = Tokenizer.from_file(a.tokenizer)
    image_id = tok.token_to_id("<image>")
    if image_id is None:
        raise SystemExit("tokenizer has no <image> token")
    ordinal = int(a.device.split(":", 1)[1]) if ":" in a.device else 0
    eng = load_model(a.model_config, a.weights, a.snapshot, ordinal, a.dtype)
    eng.set_option("kv_cache_f16", 1)
    run = engine_runner(eng, params_from_tuple, lambda v: VisionSettings(*v))
    batcher = PageBatcher(run, max_batch=a.max_batch, max_wait_ms=a.max_wait_ms)
    crop = a.crop_mode.lower() in ("1", "true", "yes", "on")
    app = create_app(batcher, tok, image_id, vision=(a.base_size, a.image_size, crop), max_new_tokens=a.max_new_tokens)
    try:
        uvicorn.run(app, host=a.host, port=a.port, log_level="info")
    finally:
        batcher.close()
        eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
