"""The drop-in boundary beyond one-image greedy calls, through the C ABI on the GPU:
  * `OcrEngine::decode` with several images per prompt and with none (model/mod.rs:2370-2603) - dsocr_decode_requests,
  * DecodeParameters.{do_sample, temperature, top_k, top_p, seed, repetition_penalty} (inference.rs:18-34, sampling.rs:34-96),
  * the streaming callback after every accepted token (model/mod.rs:1978-1982)."""
import numpy as np
import pytest
import torch

from oracle import decoder as D
from oracle import preprocess as P
from oracle import sampling as S
from tests.helpers import tiny_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from dsocr.engine import load_model

    cfg, ck, d = tiny_model("bf16")
    eng = load_model(d + "/config.json", d + "/model.safetensors", None, 0, "bf16")
    yield cfg, ck, eng, D.DecoderOracle(cfg, ck)
    eng.close()


def _image_rows(eng, page, vs):
    g, tiles, crop = eng.preprocess_gpu(page, vs)
    return eng.vision_encode_u8([g], [tiles if len(tiles) else None], [crop], vs.image_size)[0]


def test_multi_image_and_text_only_requests(setup):
    from dsocr.engine import DecodeParameters, VisionSettings

    cfg, ck, eng, oracle = setup
    vs = VisionSettings(1024, 640, True)
    pages = [P.synthetic_page(700, 1400, seed=1), P.synthetic_page(600, 500, seed=2), P.synthetic_page(1300, 640, seed=3)]
    seg = lambda *t: [int(x) for x in t]
    requests = [
        ([pages[0], pages[1]], [seg(11, 12), seg(13), seg(14, 15, 16)]),   # two <image> slots
        ([], [seg(21, 22, 23, 24, 25)]),                                      # text only: the vision tower is skipped
        ([pages[2]], [[], seg(31, 32)]),                                      # the usual single slot
        ([pages[1]], [seg(41), []]),
        ([], [seg(51, 52)]),
    ]
    params = DecodeParameters(max_new_tokens=12, no_repeat_ngram_size=20, eos_token_id=None)
    seen = {}
    outs = eng.decode_requests(requests, vs, cfg.image_token_id, params,
                               callback=lambda r, c, t: seen.setdefault(r, []).append((c, list(t))))
    with torch.no_grad():
        for r, (images, segments) in enumerate(requests):
            rows = [_image_rows(eng, im, vs) for im in images]
            ids, mask = D.build_prompt_tokens(segments, [x.shape[0] for x in rows], cfg)
            allrows = torch.from_numpy(np.concatenate(rows)) if rows else None
            ref = oracle.generate(ids, mask, allrows, 12, 20, None)
            assert outs[r].prompt_tokens == len(ids)
            assert outs[r].generated_tokens == ref, r
            # one callback per accepted token, in order, with the running list (model/mod.rs:1978-1982)
            assert [c for c, _ in seen[r]] == list(range(1, 13))
            assert all(t == ref[:c] for c, t in seen[r])


def test_slot_mismatch_error(setup):
    from dsocr.binding import DsocrError
    from dsocr.engine import DecodeParameters, VisionSettings

    cfg, ck, eng, oracle = setup
    page = P.synthetic_page(640, 640, seed=5)
    with pytest.raises(DsocrError, match=r"prompt formatting failed: prompt/image embedding mismatch: 1 slots vs 2 embeddings"):
        eng.decode_requests([([page, page], [[1], [2]])], VisionSettings(640, 640, False), cfg.image_token_id,
                            DecodeParameters(max_new_tokens=2))
    with pytest.raises(DsocrError, match=r"mismatch: 2 slots vs 0 embeddings"):
        eng.decode_requests([([], [[1], [2], [3]])], VisionSettings(640, 640, False), cfg.image_token_id,
                            DecodeParameters(max_new_tokens=2))


def _oracle_sampled(oracle, cfg, ids, mask, rows, steps, seed, **kw):
    rng = S.init_rng(seed)
    ctx = list(ids)
    kv = oracle.new_cache()
    rt = None if rows is None else torch.from_numpy(rows)
    emb = oracle.inject(oracle.embed(torch.tensor(ids)), torch.tensor(mask, dtype=torch.bool), rt)
    logits = oracle.forward(emb, 0, kv, last_only=True)[0]
    out, pos = [], len(ctx)
    for _ in range(steps):
        t = S.select_token_id(logits.numpy(), ctx, rng, **kw)
        ctx.append(t); out.append(t)
        if len(out) == steps:
            break
        logits = oracle.forward(oracle.embed(torch.tensor([t])), pos, kv)[0]
        pos += 1
    return out


@pytest.mark.parametrize("n_pages", [2, 6])  # fused small-batch step / batched step
def test_seeded_sampling_matches_oracle(setup, n_pages):
    from dsocr.engine import DecodeParameters

    cfg, ck, eng, oracle = setup
    g = torch.Generator().manual_seed(40 + n_pages)
    ids, masks, rows = [], [], []
    for p in range(n_pages):
        n_img = [9, 0, 33, 5, 17, 2][p]
        text = torch.randint(2, cfg.vocab_size - 2, (6,), generator=g).tolist()
        t, m = D.build_prompt_tokens([[], text] if n_img else [text], [n_img] if n_img else [], cfg)
        ids.append(t); masks.append(m)
        rows.append((torch.randn(n_img, cfg.hidden_size, generator=g) * 0.7).numpy() if n_img else None)
    kw = dict(do_sample=True, temperature=0.8, top_k=8, top_p=0.95, repetition_penalty=1.1, no_repeat_ngram_size=20)
    params = DecodeParameters(max_new_tokens=20, eos_token_id=None, seed=11, **kw)
    got = eng.generate_batch(ids, masks, rows, params)
    assert eng.generate_batch(ids, masks, rows, params) == got  # seeded: reproducible
    greedy = eng.generate_batch(ids, masks, rows, DecodeParameters(max_new_tokens=20, eos_token_id=None))
    assert got != greedy
    with torch.no_grad():
        want = [_oracle_sampled(oracle, cfg, ids[p], masks[p], rows[p], 20, 11, **kw) for p in range(n_pages)]
    agree = sum(int(a == b) for gp, wp in zip(got, want) for a, b in zip(gp, wp)) / (20.0 * n_pages)
    print(f"[parity] seeded sampling, {n_pages} pages: token agreement with the oracle {agree:.3f}")
    # a draw that lands within the logits tolerance of a cumulative-weight boundary may legitimately differ and the
    # sequences diverge from there; with top_k = 8 that is rare
    assert agree >= 0.9
    # unseeded calls draw from the OS entropy source
    unseeded = DecodeParameters(max_new_tokens=20, eos_token_id=None, seed=None, **kw)
    assert eng.generate_batch(ids, masks, rows, unseeded) != eng.generate_batch(ids, masks, rows, unseeded)
