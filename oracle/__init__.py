"""CPU oracle for the DeepSeek-OCR per-page forward path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the algorithm of
TimmyOVO/deepseek-ocr.rs's `crates/infer-deepseek` hot path (image -> SAM +
CLIP + projector -> DeepSeek-V2 MoE prefill/decode, optional DSQ dequant).  It
is imported only by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s
`cpu_baseline` / `--impl reference` legs, and only as the checker / baseline -
never by the product path under `deepseek-ocr.rs_b200/`.

Parity status: the reference cannot be compiled here (no cargo/rustc; the
arithmetic lives in candle 0.9.2 which is not vendored) and its golden
`baselines/` fixtures and the checkpoint are not shipped, so real-checkpoint
numerics are **parity unpinned**.  What IS pinned against the reference's own
tests or independent implementations available offline is listed in
DESIGN.md section "Oracle pinning" (integer resampler vs Pillow on downscales,
window-partition / pos-embed shape tests, DSQ container bytes vs the reader
test, Q8_0/Q4_K/Q6_K dequant vs gguf-py, SAM/CLIP vs vLLM's deepencoder.py,
AA-bicubic vs torch interpolate).

Every function cites the reference file:line it follows (paths relative to
/root/reference/).
"""
