"""Parity of the BATCHED decode step - the path bench.py times - against the f32 oracle.

Decode steps of more than 4 pages run `Engine::decoder_forward(decode_mode=true)` (csrc/engine_decode.cu): split-K tcgen05
projections, `rope_attn_decode_bulk_kernel` (RoPE + KV append + attention over the cache), `post_attn_kernel` (o_proj
reduce + RMSNorm + router + top-6 + dispatch), the stream-K expert GEMM `linear_sk_kernel` over fixed-capacity expert
segments, `combine_norm_kernel`; more than 256 pages take the unfused grouped-GEMM schedule.  Every case drives that
path through the C ABI with >= 64 decode steps (so the cached K/V of a page spans several 16 KB bulk-copy stages), for
both operand types and both KV-cache storage types:
  * teacher-forced logits of every step:  max-abs <= 2e-3 * max|logit| (f32 KV) / 5e-3 (f16 KV), cosine > 0.99999,
  * free-running greedy tokens (no-repeat-ngram 20): identical to the oracle with the f32 cache; >= 95 % with f16 KV
    (BASELINE.json target) - in practice identical on these fixtures, which the test prints.
Reference: model/mod.rs:1870-2048 (generate), block.rs:123-190, 446-804, 1215-1395.
"""
from functools import lru_cache

import numpy as np
import pytest
import torch

from oracle import decoder as D
from tests.helpers import report, tiny_model

pytestmark = pytest.mark.gpu

STEPS = 66


@pytest.fixture(scope="module", params=["bf16", "f16"])
def setup(request):
    from dsocr.engine import load_model

    dtype = request.param
    cfg, ck, d = tiny_model(dtype)
    eng = load_model(d + "/config.json", d + "/model.safetensors", None, 0, dtype)
    yield dtype, cfg, eng
    eng.close()


def _prompts(cfg, n_pages, seed):
    """Mixed prompts: text-only pages, short and long image spans (lengths 6 .. ~300 tokens)."""
    g = torch.Generator().manual_seed(seed)
    ids, masks, rows = [], [], []
    for p in range(n_pages):
        n_img = [0, 3, 17, 64, 130, 273, 41, 9][p % 8] + (p // 8) % 5
        text = torch.randint(2, cfg.vocab_size - 2, (5 + p % 4,), generator=g).tolist()
        t, m = D.build_prompt_tokens([[], text] if n_img else [text], [n_img] if n_img else [], cfg)
        ids.append(t)
        masks.append(m)
        rows.append((torch.randn(n_img, cfg.hidden_size, generator=g) * 0.7).numpy() if n_img else None)
    return ids, masks, rows


def _forced(cfg, n_pages, steps, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(2, cfg.vocab_size - 2, (steps,), generator=g).tolist() for _ in range(n_pages)]


@lru_cache(maxsize=None)
def _oracle_runs(dtype, n_pages, seed, steps, check):
    """Oracle results (independent of the KV storage mode) for the pages in `check`: forced logits + selections, free tokens."""
    cfg, ck, _ = tiny_model(dtype)
    oracle = D.DecoderOracle(cfg, ck)
    ids, masks, rows = _prompts(cfg, n_pages, seed)
    forced = _forced(cfg, n_pages, steps, seed + 1)
    out = {}
    with torch.no_grad():
        for p in check:
            rt = None if rows[p] is None else torch.from_numpy(rows[p])
            lg = []
            sel = oracle.generate(ids[p], masks[p], rt, steps, 20, None, forced=forced[p], logits_out=lg)
            free = oracle.generate(ids[p], masks[p], rt, steps, 20, None)
            out[p] = (torch.stack(lg), sel, free)
    return out


def _check_pages(n_pages):
    """Pages compared with the oracle (all pages run on the GPU; the CPU oracle costs ~20 ms per token step)."""
    if n_pages <= 6:
        return tuple(range(n_pages))
    return tuple(sorted(set(list(range(0, n_pages, max(1, n_pages // 5)))[:5] + [n_pages - 1])))


@pytest.mark.parametrize("kv", ["f32", "f16"])
@pytest.mark.parametrize("n_pages", [5, 16, 64])
def test_batched_decode_logits_and_tokens(setup, n_pages, kv):
    from dsocr.engine import DecodeParameters

    dtype, cfg, eng = setup
    seed = 100 + n_pages
    ids, masks, rows = _prompts(cfg, n_pages, seed)
    forced = _forced(cfg, n_pages, STEPS, seed + 1)
    check = _check_pages(n_pages)
    ref = _oracle_runs(dtype, n_pages, seed, STEPS, check)
    params = DecodeParameters(max_new_tokens=STEPS, no_repeat_ngram_size=20, eos_token_id=None)
    eng.set_option("kv_cache_f16", 1 if kv == "f16" else 0)
    try:
        launches0 = eng.launch_count()
        sel, logits = eng.generate_forced(ids, masks, rows, params, forced, want_logits=True)
        free = eng.generate_batch(ids, masks, rows, params)
        assert eng.launch_count() > launches0
    finally:
        eng.set_option("kv_cache_f16", 0)
    tol = 2e-3 if kv == "f32" else 5e-3
    worst, agree_forced, agree_free, n_tok = 0.0, 0, 0, 0
    for p in check:
        ref_logits, ref_sel, ref_free = ref[p]
        got = torch.from_numpy(logits[p])
        err, scale, c = report(f"{dtype} {kv}-KV {n_pages} pages, page {p} (prompt {len(ids[p])})", got, ref_logits)
        worst = max(worst, err / scale)
        assert err <= tol * scale and c > 0.99999
        agree_forced += sum(int(a == b) for a, b in zip(sel[p], ref_sel))
        agree_free += sum(int(a == b) for a, b in zip(free[p], ref_free))
        n_tok += STEPS
        assert len(free[p]) == STEPS
        if kv == "f32":
            assert sel[p] == ref_sel
            assert free[p] == ref_free
    print(f"[parity] batched decode {dtype} {kv}-KV P={n_pages}: worst rel logit err {worst:.3g}, forced argmax agreement "
          f"{agree_forced / n_tok:.4f}, free-running token agreement {agree_free / n_tok:.4f} over {n_tok} tokens")
    assert agree_forced / n_tok >= 0.95 and agree_free / n_tok >= 0.95


@pytest.mark.parametrize("n_pages", [130, 256, 300])
def test_large_decode_batches(setup, n_pages):
    """128-token expert tiles (65..256 pages) and the unfused schedule beyond 256 pages; f16 KV as in the benchmark."""
    from dsocr.engine import DecodeParameters

    dtype, cfg, eng = setup
    if dtype == "f16" and n_pages != 256:
        pytest.skip("covered with the bf16 engine")
    steps, seed = 24, 500 + n_pages
    ids, masks, rows = _prompts(cfg, n_pages, seed)
    check = _check_pages(n_pages)
    ref = _oracle_runs(dtype, n_pages, seed, steps, check)
    params = DecodeParameters(max_new_tokens=steps, no_repeat_ngram_size=20, eos_token_id=None)
    eng.set_option("kv_cache_f16", 1)
    try:
        free = eng.generate_batch(ids, masks, rows, params)
    finally:
        eng.set_option("kv_cache_f16", 0)
    agree = sum(int(a == b) for p in check for a, b in zip(free[p], ref[p][2]))
    print(f"[parity] {dtype} f16-KV P={n_pages}: free-running agreement {agree / (len(check) * steps):.4f}")
    assert agree / (len(check) * steps) >= 0.95
    eng.set_option("kv_cache_f16", 0)
    free32 = eng.generate_batch(ids, masks, rows, params)
    for p in check:
        assert free32[p] == ref[p][2]


def test_batched_eos_and_bans(setup):
    """Pages of one batch stop at different steps (EOS is not appended, model/mod.rs:2029-2033) while the others go on, with
    a small n-gram so that bans fire, and with a repetition penalty (sampling.rs:120-139) - all on the device path."""
    from dsocr.engine import DecodeParameters

    dtype, cfg, eng = setup
    cfg, ck, _ = tiny_model(dtype)
    oracle = D.DecoderOracle(cfg, ck)
    ids, masks, rows = _prompts(cfg, 8, seed=77)
    rts = [None if r is None else torch.from_numpy(r) for r in rows]
    with torch.no_grad():
        free = [oracle.generate(ids[p], masks[p], rts[p], 40, 2, None) for p in range(8)]
    eos = free[3][11]
    with torch.no_grad():
        want = [oracle.generate(ids[p], masks[p], rts[p], 40, 2, eos) for p in range(8)]
    assert len({len(w) for w in want}) > 1  # pages really stop at different steps
    got = eng.generate_batch(ids, masks, rows, DecodeParameters(40, no_repeat_ngram_size=2, eos_token_id=eos))
    assert got == want

    from oracle import sampling as S

    def oracle_penalised(p, steps, penalty):
        # same loop as DecoderOracle.generate with sampling.rs's penalty in the selection
        ctx = list(ids[p])
        kv = oracle.new_cache()
        emb = oracle.inject(oracle.embed(torch.tensor(ids[p])), torch.tensor(masks[p], dtype=torch.bool), rts[p])
        logits = oracle.forward(emb, 0, kv, last_only=True)[0]
        out, pos = [], len(ctx)
        for _ in range(steps):
            t = S.select_token_id(logits.numpy(), ctx, None, repetition_penalty=penalty, no_repeat_ngram_size=20)
            ctx.append(t); out.append(t)
            if len(out) == steps:
                break
            logits = oracle.forward(oracle.embed(torch.tensor([t])), pos, kv)[0]
            pos += 1
        return out

    with torch.no_grad():
        want_pen = [oracle_penalised(p, 24, 1.3) for p in range(6)]
    got_pen = eng.generate_batch(ids[:6], masks[:6], rows[:6], DecodeParameters(24, repetition_penalty=1.3, eos_token_id=None))
    assert got_pen == want_pen
    assert want_pen != [f[:24] for f in free[:6]]


def test_thousands_of_bans(setup):
    """the reference's banned set is an unbounded HashSet (sampling.rs:141-158).  Bigram ban with a prompt in
    which token A was followed by almost every token of the vocabulary: ~2000 bans, the selection must come from the
    few tokens left (a 64-entry ban list would pick a banned one)."""
    from dsocr.engine import DecodeParameters

    dtype, cfg, eng = setup
    if dtype != "bf16":
        pytest.skip("selection kernel is independent of the operand type")
    cfg, ck, _ = tiny_model(dtype)
    oracle = D.DecoderOracle(cfg, ck)
    a = 7
    ids, masks = [], []
    for page in range(5):  # 5 pages -> batched decode path
        spared = set(range(300 + 11 * page, 300 + 11 * page + 9)) | {a}
        text = []
        for f in range(2, cfg.vocab_size - 1):
            if f not in spared:
                text += [a, f]
        t, m = D.build_prompt_tokens([text + [a]], [], cfg)
        ids.append(t); masks.append(m)
    chk = (0, 3)
    with torch.no_grad():
        want = {p: oracle.generate(ids[p], masks[p], None, 4, 2, None) for p in chk}
        unbanned = {p: oracle.generate(ids[p], masks[p], None, 1, None, None)[0] for p in chk}
    got = eng.generate_batch(ids, masks, [None] * 5, DecodeParameters(4, no_repeat_ngram_size=2, eos_token_id=None))
    for page in range(5):
        allowed = set(range(300 + 11 * page, 300 + 11 * page + 9)) | {0, 1, a, cfg.vocab_size - 1}
        assert got[page][0] in allowed
    for p in chk:
        assert got[p] == want[p]
    assert any(unbanned[p] != want[p][0] for p in chk)  # the ban changed the selection
