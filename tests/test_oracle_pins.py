"""Pin the CPU oracle against everything reproducible offline: the reference's own self-contained
tests, Pillow, torch's antialiased interpolate, and vLLM's independent PyTorch statement of the SAM helpers."""
import ast
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from PIL import Image

from oracle import decoder as D
from oracle import preprocess as P
from oracle import vision as V
from oracle.config import tiny_config, random_checkpoint


# --- integer preprocessing ---------------------------------------------------------------------
@pytest.mark.parametrize("src,dst", [((517, 333), (200, 128)), ((1654, 2339), (724, 1024)), ((2852, 1756), (1024, 630)),
                                     ((800, 600), (640, 480)), ((641, 1281), (640, 1280))])
def test_resize_bicubic_bit_exact_with_pillow_on_downscale(src, dst):
    """vision/resample.rs restates Pillow's fixed-point bicubic; bit-identical for downscales."""
    rng = np.random.RandomState(src[0])
    img = rng.randint(0, 256, (src[1], src[0], 3), dtype=np.uint8)
    got = P.resize_bicubic(img, dst[0], dst[1])
    ref = np.asarray(Image.fromarray(img).resize(dst, Image.BICUBIC))
    assert np.array_equal(got, ref)


def test_resize_bicubic_upscale_follows_resample_rs_not_pillow():
    """round_half_towards_zero uses ceil(v+0.5) for negative v (resample.rs:18-24) -> differs from Pillow on upscales."""
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (333, 517, 3), dtype=np.uint8)
    got = P.resize_bicubic(img, 700, 500)
    ref = np.asarray(Image.fromarray(img).resize((700, 500), Image.BICUBIC))
    assert got.shape == ref.shape and not np.array_equal(got, ref)
    assert np.abs(got.astype(int) - ref).max() <= 32


def test_tile_grid_selection_and_token_counts():
    assert P.select_tile_grid(1654, 2339, 640) == (2, 3)          # A4 portrait (903 image tokens)
    assert P.select_tile_grid(2852, 1756, 640) == (3, 2)          # assets/sample_1.png
    assert P.image_token_count(1024, 640, True, (2, 3)) == 903
    assert P.image_token_count(1024, 1024, False, None) == 273     # Base-1024
    assert P.image_token_count(1024, 640, True, (1, 1)) == 273     # small image: no local tokens
    tiles, crop = P.dynamic_preprocess(np.zeros((600, 640, 3), np.uint8), 640)
    assert tiles == [] and crop == (1, 1)                          # preprocess.rs:72-80


def test_global_view_letterbox_matches_survey_probe():
    """sample_1.png is 2852x1756 -> Base mode letterboxes to 1024x630 on the 127-grey canvas."""
    img = np.full((1756, 2852, 3), 255, np.uint8)
    g = P.build_global_view(img, 1024)
    rows = np.where((g == 255).all(axis=(1, 2)))[0]
    assert len(rows) == 630 and rows[0] == 197 and (g[0] == 127).all()


def test_round_ties_to_even():
    assert [P.round_ties_to_even(v) for v in (0.5, 1.5, 2.5, -0.5, -1.5, 2.4, 2.6)] == [0.0, 2.0, 2.0, -0.0, -2.0, 2.0, 3.0]


# --- reference's own self-contained tests --------------------------------------------------------
def test_window_partition_math_matches_reference_test():
    """crates/infer-deepseek/tests/vision_sam.rs:70-81: 64x48 tokens, window 14 -> 70x56, 5x4 tiles."""
    x = torch.zeros(1, 64, 48, 8)
    win, (hp, wp) = V.window_partition(x, 14)
    assert (hp, wp) == (70, 56) and win.shape[0] == 5 * 4 and win.shape[1:3] == (14, 14)
    back = V.window_unpartition(win, 14, (hp, wp), (64, 48))
    assert back.shape == x.shape


def test_clip_position_embedding_257_to_101_tokens():
    """crates/infer-deepseek/tests/vision_clip.rs:24-34."""
    cfg = tiny_config()
    ck = {k: v for k, v in random_checkpoint(tiny_config(clip_layers=0, sam_depth=0, num_layers=0), seed=1).items()}
    clip = V.ClipOracle(cfg, ck)
    pos = clip.adapt_pos(101)
    assert pos.shape == (101, 1024)
    assert torch.equal(pos[0], clip.p["embeddings.position_embedding.weight"][0])  # cls row kept
    assert torch.equal(clip.adapt_pos(257), clip.p["embeddings.position_embedding.weight"])


# --- independent implementations available offline -----------------------------------------------
@pytest.mark.parametrize("src,dst", [(64, 40), (16, 10), (16, 24), (64, 32)])
def test_aa_bicubic_matches_torch_interpolate(src, dst):
    """vision/sam.rs:1000-1123 == F.interpolate(bicubic, antialias=True, align_corners=False) (2e-6)."""
    g = torch.Generator().manual_seed(src * 100 + dst)
    x = torch.randn(1, 8, src, src, generator=g)
    got = V.bicubic_resize_antialiased(x, dst, dst)
    ref = F.interpolate(x, size=(dst, dst), mode="bicubic", antialias=True, align_corners=False)
    assert (got - ref).abs().max().item() < 5e-6


_VLLM_DEEPENCODER = "/opt/prime-rl/.venv/lib/python3.12/site-packages/vllm/model_executor/models/deepencoder.py"


def _vllm_functions(names):
    src = open(_VLLM_DEEPENCODER).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": F, "math": math, "nn": torch.nn}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), _VLLM_DEEPENCODER, "exec"), ns)
    return ns


@pytest.mark.skipif(not os.path.exists(_VLLM_DEEPENCODER), reason="vLLM deepencoder.py not installed")
@pytest.mark.parametrize("size,table_len", [(14, 27), (64, 127), (40, 127), (32, 127)])
def test_rel_pos_matches_vllm_deepencoder(size, table_len):
    """get_rel_pos / add_decomposed_rel_pos (vision/sam.rs:1124-1247) vs vLLM's PyTorch statement of SAM."""
    ns = _vllm_functions({"get_rel_pos", "add_decomposed_rel_pos"})
    g = torch.Generator().manual_seed(size)
    rel_h = torch.randn(table_len, 64, generator=g)
    rel_w = torch.randn(table_len, 64, generator=g)
    q = torch.randn(3, size * size, 64, generator=g)
    mine_h = V.get_rel_pos(size, size, rel_h)
    ref_h = ns["get_rel_pos"](size, size, rel_h)
    assert torch.allclose(mine_h, ref_h, atol=1e-6)
    ref_bh, ref_bw = ns["add_decomposed_rel_pos"](q, rel_h, rel_w, (size, size), (size, size))
    q5 = q.reshape(3, size, size, 64)
    mine_bh = torch.einsum("bhwd,hkd->bhwk", q5, mine_h)
    mine_bw = torch.einsum("bhwd,wkd->bhwk", q5, V.get_rel_pos(size, size, rel_w))
    assert torch.allclose(mine_bh.reshape(ref_bh.shape[0], -1), ref_bh.reshape(ref_bh.shape[0], -1), atol=1e-4)
    assert torch.allclose(mine_bw.reshape(ref_bw.shape[0], -1), ref_bw.reshape(ref_bw.shape[0], -1), atol=1e-4)


@pytest.mark.skipif(not os.path.exists(_VLLM_DEEPENCODER), reason="vLLM deepencoder.py not installed")
def test_window_partition_matches_vllm_deepencoder():
    ns = _vllm_functions({"window_partition", "window_unpartition"})
    x = torch.randn(2, 40, 40, 16)
    mine, pad = V.window_partition(x, 14)
    ref, pad_ref = ns["window_partition"](x, 14)
    assert pad == tuple(pad_ref) and torch.equal(mine, ref)
    assert torch.equal(V.window_unpartition(mine, 14, pad, (40, 40)), ns["window_unpartition"](ref, 14, pad_ref, (40, 40)))


# --- decoder-side semantics ------------------------------------------------------------------------
def test_rope_matches_neox_rotate_half_and_is_norm_preserving():
    cos, sin = D.rope_tables(10000.0, 128, torch.arange(5, 9))
    assert cos.shape == (4, 128) and torch.equal(cos[:, :64], cos[:, 64:])
    x = torch.randn(10, 4, 128)
    y = D.apply_rope(x, cos, sin)
    assert torch.allclose(x.norm(dim=-1), y.norm(dim=-1), atol=1e-4)
    assert torch.allclose(D.apply_rope(x, *D.rope_tables(10000.0, 128, torch.zeros(4))), x)


def test_banned_ngram_tokens_semantics():
    """crates/core/src/sampling.rs:141-158."""
    seq = [1, 2, 3, 9, 1, 2, 3, 7, 1, 2]
    assert D.banned_ngram_tokens(seq, 3) == {3}
    assert D.banned_ngram_tokens(seq + [3], 4) == {9, 7}
    assert D.banned_ngram_tokens([1, 2], 20) == set()
    assert D.banned_ngram_tokens(seq, 1) == set()


def test_first_index_argmax_and_ban():
    logits = torch.tensor([0.1, 3.0, 3.0, -1.0])
    assert D.select_token_greedy(logits, [5], 20) == 1                     # torch.argmax tie-break: first index
    assert D.select_token_greedy(logits, [0, 1, 0], 2) == 2                # token 1 banned (bigram 0,1 seen)


def test_moe_router_tie_break_lowest_index():
    cfg = tiny_config(num_layers=2, n_routed_experts=16)
    ck = random_checkpoint(cfg, seed=3)
    ck["model.layers.1.mlp.gate.weight"] = torch.zeros_like(ck["model.layers.1.mlp.gate.weight"])  # all scores tie
    taps = {}
    D.DecoderOracle(cfg, ck).moe(torch.randn(3, cfg.hidden_size), 1, taps)
    assert taps["topk_idx"][0].tolist() == [list(range(cfg.num_experts_per_tok))] * 3


def test_build_prompt_tokens_mismatch_message():
    cfg = tiny_config()
    with pytest.raises(AssertionError, match="prompt/image embedding mismatch"):
        D.build_prompt_tokens([[1, 2]], [5], cfg)
    ids, mask = D.build_prompt_tokens([[], [7, 8]], [3], cfg)
    assert ids == [0, cfg.image_token_id, cfg.image_token_id, cfg.image_token_id, 7, 8] and mask == [0, 1, 1, 1, 0, 0]


# --- DSQ: container + block formats --------------------------------------------------------------------
def test_dsq_dequant_and_q8_quantiser_match_gguf():
    """ggml block semantics pinned against gguf-py: dequant for all three formats, quantiser for Q8_0."""
    import gguf
    from gguf import quants

    from oracle import dsq

    rng = np.random.RandomState(0)
    w = (rng.randn(8, 512) * 0.02).astype(np.float32)
    assert quants.quantize(w, gguf.GGMLQuantizationType.Q8_0).tobytes() == dsq.quantize_q8_0(w)
    for dt, gg in ((dsq.Q8_0, gguf.GGMLQuantizationType.Q8_0), (dsq.Q4K, gguf.GGMLQuantizationType.Q4_K),
                   (dsq.Q6K, gguf.GGMLQuantizationType.Q6_K)):
        q = dsq._QUANT[dt](w)
        assert len(q) == 8 * 512 // dsq.BLOCK[dt] * dsq.BLOCK_BYTES[dt]
        mine = dsq.dequantize(q, dt, 8, 512)
        ref = quants.dequantize(np.frombuffer(q, dtype=np.uint8).reshape(8, -1), gg)
        assert np.array_equal(mine, ref)
        assert np.abs(mine - w).max() < 0.06 * np.abs(w).max()
    # arbitrary (random) valid K-quant blocks, not only the ones our quantiser emits
    for dt, gg in ((dsq.Q4K, gguf.GGMLQuantizationType.Q4_K), (dsq.Q6K, gguf.GGMLQuantizationType.Q6_K)):
        nb = dsq.BLOCK_BYTES[dt]
        raw = rng.randint(0, 256, (4, 3, nb), dtype=np.uint8)
        half = np.array([0.01], dtype=np.float16).view(np.uint8)
        if dt == dsq.Q4K:
            raw[..., 0:2] = half; raw[..., 2:4] = half
        else:
            raw[..., 208:210] = half
        mine = dsq.dequantize(raw.tobytes(), dt, 4, 768)
        ref = quants.dequantize(raw.reshape(4, -1), gg)
        assert np.array_equal(mine, ref)


def test_dsq_container_layout_matches_reader_test(tmp_path):
    """Byte layout of crates/dsq/tests/reader.rs:9-66 (build_snapshot_bytes): magic, version, 3 strings, dtype, block,
    count, record {name, out, in, dtype, q_offset, q_len, bias_offset, bias_len, bias_dtype}, payload, bias."""
    import struct

    from oracle import dsq

    q = bytes(range(34)) * 2 * 4        # out_dim 4, in_dim 64 -> 2 blocks per row
    bias = struct.pack("<4f", 1, 2, 3, 4)
    path = str(tmp_path / "t.dsq")
    dsq.write_snapshot(path, dsq.Q8_0, [("layer.weight", 4, 64, dsq.Q8_0, q, bias)], model_id="model-id", backend="CPU",
                       candle_version="candle-test")
    raw = open(path, "rb").read()

    def ws(s):
        return struct.pack("<I", len(s)) + s.encode()
    head = b"DSQSNAP" + struct.pack("<I", 1) + ws("candle-test") + ws("model-id") + ws("CPU") + struct.pack("<III", 8, 32, 1)
    rec_size = 4 + len("layer.weight") + 12 + 32 + 4
    q_off = len(head) + rec_size
    rec = ws("layer.weight") + struct.pack("<III", 4, 64, 8) + struct.pack("<QQQQ", q_off, len(q), q_off + len(q), 16) + struct.pack("<I", 4)
    assert raw == head + rec + q + bias
    hdr, recs, data = dsq.read_snapshot(path)
    r = recs["layer.weight"]
    assert (r.out_dim, r.in_dim, r.q_dtype, r.q_offset, r.q_len, r.bias_len) == (4, 64, 8, q_off, len(q), 16)
    bad = bytearray(raw); bad[0:7] = b"BADSNAP"
    (tmp_path / "bad.dsq").write_bytes(bytes(bad))
    with pytest.raises(ValueError, match="magic"):
        dsq.read_snapshot(str(tmp_path / "bad.dsq"))
    bad = bytearray(raw); bad[7:11] = struct.pack("<I", 2)
    (tmp_path / "bad2.dsq").write_bytes(bytes(bad))
    with pytest.raises(ValueError, match="version"):
        dsq.read_snapshot(str(tmp_path / "bad2.dsq"))


def test_dsq_dtype_assignment_follows_exporter():
    from oracle import dsq

    assert dsq.choose_dtype("model.layers.1.self_attn.q_proj.weight", 1280, dsq.Q4K) == dsq.Q4K
    assert dsq.choose_dtype("model.layers.1.mlp.experts.3.down_proj.weight", 896, dsq.Q4K) == dsq.Q8_0
    assert dsq.choose_dtype("model.layers.0.mlp.down_proj.weight", 6848, dsq.Q6K) == dsq.Q8_0
    assert dsq.choose_dtype("lm_head.weight", 1280, dsq.Q4K) == dsq.Q8_0
    assert dsq.choose_dtype("lm_head.weight", 1280, dsq.Q8_0) == dsq.Q8_0
    assert dsq.choose_dtype("x", 1792, dsq.Q6K) == dsq.Q6K


# --- KV-cache bookkeeping and padding mask: the reference's own unit tests ------------------------------
def _chunk(batch, heads, seq, dim):
    from oracle.cache import KvCacheChunk

    return KvCacheChunk(torch.zeros(batch, heads, dim, seq), torch.zeros(batch, heads, seq, dim))


def test_lengths_to_padding_mask_builds_expected():
    """crates/infer-deepseek/tests/transformer_block.rs:59-67."""
    from oracle.cache import lengths_to_padding_mask

    m = lengths_to_padding_mask([2, 4], 4)
    assert m[0].tolist() == [1.0, 1.0, 0.0, 0.0] and m[1].tolist() == [1.0, 1.0, 1.0, 1.0]
    with pytest.raises(ValueError, match="exceeds sequence dimension"):
        lengths_to_padding_mask([5], 4)


def test_layer_cache_auto_resizes_and_rejects_incompatible_dimensions():
    """transformer_cache.rs:20-47."""
    from oracle.cache import LayerKvCache

    cache = LayerKvCache()
    cache.append_chunk(1, _chunk(1, 2, 3, 4))
    assert len(cache) == 2 and cache.get(0) is None and cache.get(1).seq_len() == 3 and cache.seq_len() == 3
    cache = LayerKvCache(1)
    cache.append_chunk(0, _chunk(1, 2, 3, 4))
    with pytest.raises(ValueError, match="chunk heads"):
        cache.append_chunk(0, _chunk(1, 3, 1, 4))


def test_dynamic_cache_tracks_sequence_growth_and_guard_clears():
    """transformer_cache.rs:49-101."""
    from oracle.cache import DynamicCache

    cache = DynamicCache(3)
    cache.append(0, _chunk(1, 2, 3, 4))
    cache.append(1, _chunk(1, 2, 3, 4))
    assert cache.seq_len() == 3
    cache.append(0, _chunk(1, 2, 1, 4))
    assert cache.seq_len() == 4
    with pytest.raises(ValueError, match="seq_len decreased"):
        cache.append(2, _chunk(1, 2, 2, 4))
    cache.append(1, _chunk(1, 2, 1, 4))
    assert cache.seq_len() == 4
    assert cache.get(0).key_view().shape == (1, 2, 4, 4) and cache.get(0).value_view().shape == (1, 2, 4, 4)
    cache = DynamicCache(1)
    flag = []
    with cache.prompt_guard(reset=lambda: flag.append(True)) as c:
        c.append(0, _chunk(1, 2, 3, 4))
        assert c.seq_len() == 3
    assert cache.seq_len() is None and all(e is None for e in cache.layers.entries) and flag == [True]
