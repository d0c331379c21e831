"""Token selection behind the C ABI when `do_sample` is set (csrc/sampler.cpp) against the oracle's restatement of
crates/core/src/sampling.rs (oracle/sampling.py), plus known answers for the ChaCha core of rand 0.8's StdRng.
Runs without a GPU: the sampling branch is host code (every step's logits rows are copied back for it)."""
import ctypes as C
import math
import struct

import numpy as np
import pytest

from dsocr.binding import check, lib
from dsocr.engine import DecodeParameters
from oracle import sampling as S

# zero key, zero counter / stream: first 64 keystream bytes
KAT = {
    # RFC 7539 2.3.2 layout with zero counter/nonce == rand_chacha `test_chacha_true_values_a` (0xade0b876, 0x903df1a0, ...)
    20: "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586",
    # draft-strombergson-chacha-test-vectors TC1, 256-bit key, 12 and 8 rounds
    12: "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be",
    8: "3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42",
}


def _lib_words(key: bytes, rounds: int, n: int):
    out = (C.c_uint32 * n)()
    check(lib().dsocr_test_chacha_words((C.c_uint8 * 32)(*key), rounds, n, out), "chacha")
    return list(out)


@pytest.mark.parametrize("rounds", [20, 12, 8])
def test_chacha_known_answers(rounds):
    want = list(struct.unpack("<16I", bytes.fromhex(KAT[rounds])))
    assert S.chacha_block([0] * 8, 0, rounds) == want
    assert _lib_words(bytes(32), rounds, 16) == want


def test_block_rng_order_and_counter():
    key = bytes(range(32))
    words = _lib_words(key, 12, 200)
    rng = S.StdRng(key)
    assert words == [rng.next_u32() for _ in range(200)]
    # blocks are consecutive counters
    kw = list(struct.unpack("<8I", key))
    assert words[64:80] == S.chacha_block(kw, 4, 12)
    assert words[16:32] == S.chacha_block(kw, 1, 12)


@pytest.mark.parametrize("seed", [0, 1, 42, 2 ** 63 + 12345, 2 ** 64 - 1])
def test_seed_from_u64_draws(seed):
    n = 70  # crosses a 64-word buffer refill
    out = (C.c_uint64 * n)()
    check(lib().dsocr_test_stdrng_u64(C.c_uint64(seed), n, out), "stdrng")
    rng = S.StdRng.seed_from_u64(seed)
    assert list(out) == [rng.next_u64() for _ in range(n)]


def test_next_u64_straddles_refill():
    rng = S.StdRng(bytes(32))
    first = [rng.next_u32() for _ in range(63)]
    v = rng.next_u64()  # last word of this buffer (low) + first word of the next (high)
    ref = S.StdRng(bytes(32))
    words = [ref.next_u32() for _ in range(65)]
    assert first == words[:63] and v == (words[64] << 32) | words[63]


def _select_lib(logits, params: DecodeParameters, context):
    steps, V = logits.shape
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    ctx = np.asarray(context, dtype=np.int64)
    out = (C.c_int64 * steps)()
    p = params.c()
    check(lib().dsocr_test_select_tokens(lg.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(V), steps, C.byref(p),
                                         ctx.ctypes.data_as(C.POINTER(C.c_int64)), C.c_size_t(len(ctx)), out), "select")
    return list(out)


def _select_oracle(logits, params: DecodeParameters, context):
    rng = S.init_rng(params.seed)
    ctx = list(context)
    out = []
    for row in logits:
        t = S.select_token_id(row, ctx, rng, do_sample=params.do_sample, temperature=params.temperature, top_p=params.top_p,
                              top_k=params.top_k, repetition_penalty=params.repetition_penalty,
                              no_repeat_ngram_size=params.no_repeat_ngram_size)
        out.append(t)
        ctx.append(t)
    return out


@pytest.mark.parametrize("case", [
    dict(do_sample=True, temperature=1.0, top_p=None, top_k=None),
    dict(do_sample=True, temperature=0.7, top_p=0.9, top_k=None),
    dict(do_sample=True, temperature=1.3, top_p=None, top_k=5),
    dict(do_sample=True, temperature=0.5, top_p=0.6, top_k=12, repetition_penalty=1.3),
    dict(do_sample=True, temperature=1.0, top_p=1.0, top_k=None, no_repeat_ngram_size=2),
    dict(do_sample=True, temperature=0.0, top_p=0.5, top_k=3),            # temperature 0 -> argmax (sampling.rs:62)
    dict(do_sample=False, temperature=1.0, repetition_penalty=1.5, no_repeat_ngram_size=3),
])
def test_select_token_matches_oracle(case):
    g = np.random.default_rng(7)
    V, steps = 97, 60
    logits = (g.standard_normal((steps, V)) * 2.0).astype(np.float32)
    context = g.integers(0, V, 30).tolist()
    params = DecodeParameters(max_new_tokens=steps, eos_token_id=None, seed=1234,
                              no_repeat_ngram_size=case.pop("no_repeat_ngram_size", None), **case)
    got = _select_lib(logits, params, context)
    want = _select_oracle(logits, params, context)
    assert got == want
    if params.do_sample and params.temperature > 0:
        assert len(set(got)) > 5  # really samples
        assert _select_lib(logits, params, context) == got  # seeded -> reproducible
        other = DecodeParameters(**{**params.__dict__, "seed": 99})
        assert _select_lib(logits, other, context) != got


def test_reference_semantics_unit_cases():
    rng = S.StdRng.seed_from_u64(0)
    # first-index argmax, non-finite skipped (sampling.rs:104-118)
    assert S.select_token_id(np.array([1.0, 3.0, 3.0, np.nan, np.inf], np.float32), [], rng) == 1
    # repetition penalty: positive scores divided, others multiplied, each distinct token once (:120-139)
    lg = np.array([2.0, 1.9, -1.0, 0.5], np.float32)
    assert S.select_token_id(lg, [0, 0, 0], rng, repetition_penalty=1.2) == 1
    sc = lg.copy(); S.apply_repetition_penalty(sc, [0, 2, 0, 7, -1], 2.0)
    assert sc.tolist() == [1.0, np.float32(1.9), -2.0, 0.5]
    # n-gram ban (:141-158) and the fall-back when every finite logit is banned (:58-60)
    assert S.banned_ngram_tokens([5, 6, 7, 5, 6], 3) == {7}
    assert S.select_token_id(np.array([0.0, 9.0], np.float32), [1, 1], rng, no_repeat_ngram_size=2) == 0
    assert S.select_token_id(np.array([-np.inf, 9.0], np.float32), [1, 1], rng, no_repeat_ngram_size=2) == 1
    # top-k keeps the k largest, stable for ties (:160-174); top-p keeps the smallest prefix whose mass exceeds p (:176-224)
    l = [1.0, 3.0, 3.0, 2.0]; S.apply_top_k(l, 2); assert l == [-math.inf, 3.0, 3.0, -math.inf]
    l = [math.log(0.5), math.log(0.3), math.log(0.2)]; S.apply_top_p(l, 0.6)
    assert l[2] == -math.inf and l[0] > -math.inf and l[1] > -math.inf
    l = [math.log(0.5), math.log(0.3), math.log(0.2)]; S.apply_top_p(l, 0.4); assert l[1] == l[2] == -math.inf
    # a single surviving candidate is returned whatever the draw
    assert S.select_token_id(np.array([0.0, 5.0, 1.0], np.float32), [], rng, do_sample=True, temperature=1.0, top_k=1) == 1


def test_weighted_index_distribution():
    """The draw follows the softmax distribution (WeightedIndex over exp(l - max))."""
    logits = np.log(np.array([[0.1, 0.2, 0.3, 0.4]], np.float32)).repeat(4000, 0)
    params = DecodeParameters(max_new_tokens=1, do_sample=True, temperature=1.0, top_p=None, eos_token_id=None, seed=5,
                              no_repeat_ngram_size=None)
    got = np.bincount(_select_lib(logits, params, []), minlength=4) / 4000.0
    assert np.abs(got - np.array([0.1, 0.2, 0.3, 0.4])).max() < 0.03
