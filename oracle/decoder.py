"""f32 CPU restatement of the DeepSeek-V2 MoE decoder path.  Test infrastructure only.

Follows:
  * crates/infer-deepseek/src/transformer/rope.rs:172-207   (tables)
  * crates/infer-deepseek/src/transformer/block.rs:24-29, 446-804, 1166-1213, 1215-1395, 1403-1471, 1504-1561
  * crates/infer-deepseek/src/transformer/decoder.rs:62-196, transformer/model.rs:151-278
  * crates/infer-deepseek/src/model/mod.rs:1760-1857 (inject), :1870-2048 (generate)
  * crates/core/src/sampling.rs:34-158 (greedy + no-repeat-ngram, first-index argmax)
With `--dtype f16|bf16` the reference still computes the whole decoder in f32 from the
low-precision *stored* weights, which is what this does.
Liberty (kinder than the reference): a preallocated KV cache instead of per-step Tensor::cat.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from .config import OcrConfig


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """candle rms_norm_slow as used by block.rs:24-29: x / sqrt(mean(x^2) + eps) * w, f32."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def rope_tables(theta: float, dim: int, positions: torch.Tensor):
    """rope.rs:172-207: inv_freq_i = 1/theta^(2i/dim) in f32; cos/sin duplicated halves."""
    half = dim // 2
    expo = (torch.arange(half, dtype=torch.float32) * 2.0) / float(dim)
    inv = 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float32), expo)
    ang = positions.to(torch.float32).unsqueeze(1) * inv.unsqueeze(0)
    return torch.cat([ang.cos(), ang.cos()], 1), torch.cat([ang.sin(), ang.sin()], 1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """block.rs:1403-1471 with use_mla=false (NeoX rotate-half).  x [h, s, d], cos/sin [s, d]."""
    d = x.shape[-1]
    rot = torch.cat([-x[..., d // 2:], x[..., : d // 2]], -1)
    return x * cos + rot * sin


def banned_ngram_tokens(seq: Sequence[int], ngram: int) -> set:
    """core/src/sampling.rs:141-158."""
    banned = set()
    if ngram <= 1 or len(seq) < ngram - 1:
        return banned
    prefix = tuple(seq[len(seq) - (ngram - 1):])
    for i in range(len(seq) - ngram + 1):
        if tuple(seq[i: i + ngram - 1]) == prefix:
            banned.add(seq[i + ngram - 1])
    return banned


def select_token_greedy(logits: torch.Tensor, context: Sequence[int], no_repeat_ngram: Optional[int]) -> int:
    """core/src/sampling.rs:34-118 for do_sample=false, repetition_penalty=1."""
    filt = logits.clone()
    if no_repeat_ngram and no_repeat_ngram > 1:
        for t in banned_ngram_tokens(context, no_repeat_ngram):
            if 0 <= t < filt.numel():
                filt[t] = float("-inf")
    if not torch.isfinite(filt).any():
        filt = logits
    return int(torch.argmax(torch.where(torch.isfinite(filt), filt, torch.full_like(filt, float("-inf")))))


class DecoderOracle:
    def __init__(self, cfg: OcrConfig, ckpt: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.ckpt = ckpt  # stored dtype; upcast lazily per use to bound memory
        self._cache: Dict[str, torch.Tensor] = {}

    def w(self, name: str) -> torch.Tensor:
        t = self._cache.get(name)
        if t is None:
            t = self.ckpt[name].to(torch.float32)
            self._cache[name] = t
        return t

    # ---- pieces --------------------------------------------------------------------------
    def embed(self, ids: torch.Tensor) -> torch.Tensor:
        return self.ckpt["model.embed_tokens.weight"][ids].to(torch.float32)

    def inject(self, emb: torch.Tensor, mask: torch.Tensor, img_rows: Optional[torch.Tensor]) -> torch.Tensor:
        """model/mod.rs:1760-1857: masked rows <- image rows in order."""
        if img_rows is None or not mask.any():
            return emb
        out = emb.clone()
        pos = mask.nonzero().flatten()
        assert pos.numel() == img_rows.shape[0], "image embeddings provide %d tokens but mask requires %d" % (
            img_rows.shape[0], pos.numel())
        out[pos] = img_rows.to(torch.float32)
        return out

    def mlp(self, x: torch.Tensor, prefix: str) -> torch.Tensor:
        """block.rs:1166-1177."""
        g = F.linear(x, self.w(prefix + "gate_proj.weight"))
        u = F.linear(x, self.w(prefix + "up_proj.weight"))
        return F.linear(F.silu(g) * u, self.w(prefix + "down_proj.weight"))

    def moe(self, x: torch.Tensor, layer: int, taps=None) -> torch.Tensor:
        """block.rs:1215-1395.  x [T, H] f32."""
        cfg = self.cfg
        p = f"model.layers.{layer}.mlp."
        logits = F.linear(x, self.w(p + "gate.weight"))
        scores = torch.softmax(logits, -1)
        # stable descending sort => lowest index wins ties (CPU sort_last_dim, block.rs:1282)
        sorted_scores, sorted_idx = torch.sort(scores, dim=-1, descending=True, stable=True)
        k = cfg.num_experts_per_tok
        topw, topi = sorted_scores[:, :k], sorted_idx[:, :k]
        if taps is not None:
            taps.setdefault("topk_idx", []).append(topi.clone())
            taps.setdefault("topk_w", []).append(topw.clone())
        flat_i = topi.reshape(-1)
        ys = torch.zeros(flat_i.numel(), x.shape[1], dtype=torch.float32)
        for e in range(cfg.n_routed_experts):  # per-expert loop, block.rs:1329-1355
            sel = (flat_i == e).nonzero().flatten()
            if sel.numel() == 0:
                continue
            ys[sel] = self.mlp(x[sel // k], f"{p}experts.{e}.")
        # scatter back + weighted sum over k in slot order, block.rs:1363-1381
        out = (ys.reshape(x.shape[0], k, -1) * topw.unsqueeze(-1)).sum(1)
        return out + self.mlp(x, p + "shared_experts.")

    def layer(self, i: int, x: torch.Tensor, pos0: int, kv, taps=None) -> torch.Tensor:
        """block.rs:123-190 (+ attention_forward :446-804).  x [T, H]; kv = (K [L][h,S,d], V) lists."""
        cfg = self.cfg
        p = f"model.layers.{i}."
        t = x.shape[0]
        nh, hd = cfg.num_heads, cfg.head_dim
        n = rms_norm(x, self.w(p + "input_layernorm.weight"), cfg.rms_norm_eps)
        q = F.linear(n, self.w(p + "self_attn.q_proj.weight")).reshape(t, nh, hd).transpose(0, 1)
        k = F.linear(n, self.w(p + "self_attn.k_proj.weight")).reshape(t, nh, hd).transpose(0, 1)
        v = F.linear(n, self.w(p + "self_attn.v_proj.weight")).reshape(t, nh, hd).transpose(0, 1)
        cos, sin = rope_tables(cfg.rope_theta, hd, torch.arange(pos0, pos0 + t))
        q = apply_rope(q, cos, sin)
        k = apply_rope(k, cos, sin)
        if kv[0][i] is None:
            kk, vv = k, v
        else:
            kk = torch.cat([kv[0][i], k], 1)
            vv = torch.cat([kv[1][i], v], 1)
        kv[0][i], kv[1][i] = kk, vv
        scores = (q @ kk.transpose(1, 2)) / math.sqrt(hd)
        if pos0 == 0 and t > 1:
            causal = torch.triu(torch.ones(t, t, dtype=torch.bool), diagonal=1)
            scores = scores + causal.to(torch.float32) * (-1e9)  # additive bias block.rs:1497-1526
        attn = torch.softmax(scores, -1)
        ctx = (attn @ vv).transpose(0, 1).reshape(t, nh * hd)
        x = x + F.linear(ctx, self.w(p + "self_attn.o_proj.weight"))
        n2 = rms_norm(x, self.w(p + "post_attention_layernorm.weight"), cfg.rms_norm_eps)
        if i < cfg.first_k_dense_replace:
            m = self.mlp(n2, p + "mlp.")
        else:
            m = self.moe(n2, i, taps)
        return x + m

    def forward(self, emb: torch.Tensor, pos0: int, kv, last_only: bool = False, taps=None) -> torch.Tensor:
        """transformer/model.rs:151-278: layers -> final RMSNorm -> lm_head.  Returns logits [T, V]
        (the reference computes all rows; `last_only` is an oracle-side shortcut with identical values)."""
        x = emb
        for i in range(self.cfg.num_layers):
            x = self.layer(i, x, pos0, kv, taps)
            if taps is not None:
                taps.setdefault("hidden", []).append(x.clone())
        if last_only:
            x = x[-1:]
        n = rms_norm(x, self.w("model.norm.weight"), self.cfg.rms_norm_eps)
        return F.linear(n, self.w("lm_head.weight"))

    def new_cache(self):
        return ([None] * self.cfg.num_layers, [None] * self.cfg.num_layers)

    # ---- generation ----------------------------------------------------------------------
    def generate(self, input_ids: Sequence[int], mask: Sequence[int], img_rows: Optional[torch.Tensor],
                 max_new_tokens: int, no_repeat_ngram: Optional[int] = 20, eos: Optional[int] = None,
                 callback: Optional[Callable[[int, List[int]], None]] = None,
                 forced: Optional[Sequence[int]] = None, logits_out: Optional[list] = None) -> List[int]:
        """model/mod.rs:1870-2048.  `forced` = teacher forcing (tokens fed instead of the selected
        ones; selections are still returned) for per-step argmax-agreement measurements."""
        if max_new_tokens == 0:
            return []
        ids = torch.tensor(list(input_ids), dtype=torch.long)
        m = torch.tensor(list(mask), dtype=torch.bool)
        emb = self.inject(self.embed(ids), m, img_rows)
        kv = self.new_cache()
        context = list(input_ids)
        logits = self.forward(emb, 0, kv, last_only=True)[0]
        if logits_out is not None:
            logits_out.append(logits.clone())
        cur = select_token_greedy(logits, context, no_repeat_ngram)
        if eos is not None and cur == eos:
            return []
        generated: List[int] = []
        selected: List[int] = []
        pos = len(context)
        for step in range(max_new_tokens):
            selected.append(cur)
            feed = cur if forced is None else int(forced[step])
            context.append(feed)
            generated.append(feed)
            if callback is not None:
                callback(len(generated), generated)
            if step + 1 == max_new_tokens:
                break
            e = self.embed(torch.tensor([feed]))
            logits = self.forward(e, pos, kv)[0]
            pos += 1
            if logits_out is not None:
                logits_out.append(logits.clone())
            cur = select_token_greedy(logits, context, no_repeat_ngram)
            if eos is not None and cur == eos and forced is None:
                break
        return generated if forced is None else selected


def build_prompt_tokens(text_segments_ids: Sequence[Sequence[int]], n_img_tokens: Sequence[int], cfg: OcrConfig):
    """model/mod.rs:2536-2603 with the tokenizer factored out: BOS(0) + ids(seg0) + <image>*n0 + ids(seg1)...
    `text_segments_ids[i]` are the already-tokenised text segments around the image slots."""
    assert len(text_segments_ids) - 1 == len(n_img_tokens), (
        "prompt/image embedding mismatch: %d slots vs %d embeddings" % (len(text_segments_ids) - 1, len(n_img_tokens)))
    toks = [cfg.bos_token_id]
    mask = [0]
    for i, seg in enumerate(text_segments_ids):
        toks.extend(int(t) for t in seg)
        mask.extend([0] * len(seg))
        if i < len(n_img_tokens):
            toks.extend([cfg.image_token_id] * n_img_tokens[i])
            mask.extend([1] * n_img_tokens[i])
    return toks, mask
