#!/usr/bin/env python
"""End-to-end pins of the oracle's towers against implementations that are INDEPENDENT of this repo
(VERDICT r1 next #9).  Run in the build container; writes tests/golden/towers_tiny.npz, which
tests/test_oracle_e2e_pins_cpu.py holds the oracle to on any box (the generators need not exist there).

  SAM   : vLLM's PyTorch port of DeepSeek-OCR's `deepencoder` SAM ViT (ImageEncoderViT / Block / RelPosAttention /
          window_partition / add_decomposed_rel_pos / neck / net_2 / net_3, vllm/model_executor/models/deepencoder.py),
          executed from its source file with vLLM's runtime imports replaced by torch equivalents
          (Conv2dLayer -> nn.Conv2d, PluggableLayer -> nn.Module); no line of its math is changed.
  CLIP  : Hugging Face transformers' CLIPEncoder (quick_gelu, separate q/k/v projections, eps 1e-5) fed with
          [cls ; SAM tokens] + vLLM's DeepCLIPVisionEmbeddings.get_abs_pos (pos-embed 257 -> g*g+1) and pre_layrnorm.
  DEC   : Hugging Face LlamaForCausalLM (MHA 10 x 128, NeoX rotate-half RoPE theta 10000, RMSNorm 1e-6, SwiGLU) whose MLP
          of every layer >= first_k_dense_replace is replaced by transformers' DeepseekV2Moe (softmax router, greedy
          top-6, norm_topk_prob off, routed_scaling_factor 1, shared experts) - the architecture the DeepSeek-OCR
          checkpoint declares (use_mla = false).  Prefill logits of the last row + teacher-forced decode steps through
          HF's DynamicCache.

The model is the seeded tiny random-init checkpoint of oracle/config.py (values rounded to bf16, computed in f32).
Usage: python tests/golden/make_golden_towers.py [--full]   (--full additionally compares at the full depth and prints
the deviations; nothing is stored for it)."""
from __future__ import annotations

import argparse
import hashlib
import importlib.util
import math
import re
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import config as OC, decoder as D, preprocess as P, vision as V  # noqa: E402


def load_vllm_deepencoder():
    spec = importlib.util.find_spec("vllm")
    path = Path(spec.submodule_search_locations[0]) / "model_executor" / "models" / "deepencoder.py"
    src = path.read_text()
    src = re.sub(r"^from vllm\.[^\n]*$", "", src, flags=re.M)
    src = re.sub(r"^from \.clip import[^\n]*$", "", src, flags=re.M)

    class PluggableLayer(nn.Module):
        @staticmethod
        def register(_name):
            return lambda cls: cls

    ns = {"PluggableLayer": PluggableLayer, "Conv2dLayer": nn.Conv2d, "MMEncoderAttention": None, "QuantizationConfig": None,
          "default_weight_loader": None, "CLIPEncoder": object, "CLIPVisionEmbeddings": nn.Module, "__name__": "vllm_deepencoder"}
    exec(compile(src, str(path), "exec"), ns)
    return ns


def sam_reference(ns, cfg, ck):
    m = ns["_build_sam"](cfg.sam_dim, cfg.sam_depth, cfg.sam_heads, list(cfg.sam_global_idx))
    sd = {k[len("model.sam_model."):]: v.float() for k, v in ck.items() if k.startswith("model.sam_model.")}
    missing, unexpected = m.load_state_dict(sd, strict=True), None
    return m.eval()


def clip_reference(ns, cfg, ck, sam_out):
    from transformers import CLIPVisionConfig
    from transformers.models.clip.modeling_clip import CLIPEncoder

    hc = CLIPVisionConfig(hidden_size=cfg.clip_dim, intermediate_size=4 * cfg.clip_dim, num_hidden_layers=cfg.clip_layers,
                          num_attention_heads=cfg.clip_heads, hidden_act="quick_gelu", layer_norm_eps=1e-5,
                          image_size=cfg.clip_image_size, patch_size=cfg.clip_patch, attention_dropout=0.0)
    hc._attn_implementation = "eager"
    enc = CLIPEncoder(hc).eval()
    pre = "model.vision_model."
    sd = {}
    for i in range(cfg.clip_layers):
        p = f"{pre}transformer.layers.{i}."
        qkv_w, qkv_b = ck[p + "self_attn.qkv_proj.weight"].float(), ck[p + "self_attn.qkv_proj.bias"].float()
        for j, n in enumerate("qkv"):
            sd[f"layers.{i}.self_attn.{n}_proj.weight"] = qkv_w[j * cfg.clip_dim:(j + 1) * cfg.clip_dim]
            sd[f"layers.{i}.self_attn.{n}_proj.bias"] = qkv_b[j * cfg.clip_dim:(j + 1) * cfg.clip_dim]
        sd[f"layers.{i}.self_attn.out_proj.weight"] = ck[p + "self_attn.out_proj.weight"].float()
        sd[f"layers.{i}.self_attn.out_proj.bias"] = ck[p + "self_attn.out_proj.bias"].float()
        for a, b in (("layer_norm1", "layer_norm1"), ("layer_norm2", "layer_norm2"), ("mlp.fc1", "mlp.fc1"), ("mlp.fc2", "mlp.fc2")):
            sd[f"layers.{i}.{a}.weight"] = ck[p + b + ".weight"].float()
            sd[f"layers.{i}.{a}.bias"] = ck[p + b + ".bias"].float()
    enc.load_state_dict(sd, strict=True)
    b, c, gh, gw = sam_out.shape
    patches = sam_out.flatten(2).transpose(1, 2)
    cls = ck[pre + "embeddings.class_embedding"].float().reshape(1, 1, c).expand(b, 1, c)
    emb = torch.cat([cls, patches], 1)
    table = ck[pre + "embeddings.position_embedding.weight"].float().unsqueeze(0)
    pos = ns["DeepCLIPVisionEmbeddings"].get_abs_pos(None, table, emb.size(1))
    emb = emb + pos
    x = torch.nn.functional.layer_norm(emb, (c,), ck[pre + "pre_layrnorm.weight"].float(), ck[pre + "pre_layrnorm.bias"].float(), 1e-5)
    return enc(inputs_embeds=x).last_hidden_state


def decoder_reference(cfg, ck):
    from transformers import LlamaConfig, LlamaForCausalLM
    from transformers.models.deepseek_v2.configuration_deepseek_v2 import DeepseekV2Config
    from transformers.models.deepseek_v2.modeling_deepseek_v2 import DeepseekV2Moe

    lc = LlamaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                     num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads, num_key_value_heads=cfg.num_heads,
                     head_dim=cfg.head_dim, hidden_act="silu", max_position_embeddings=8192, rms_norm_eps=cfg.rms_norm_eps,
                     rope_theta=cfg.rope_theta, attention_bias=False, mlp_bias=False, tie_word_embeddings=False)
    lc._attn_implementation = "eager"
    model = LlamaForCausalLM(lc).eval()
    dc = DeepseekV2Config(hidden_size=cfg.hidden_size, moe_intermediate_size=cfg.moe_intermediate_size,
                          n_routed_experts=cfg.n_routed_experts, n_shared_experts=cfg.n_shared_experts,
                          num_experts_per_tok=cfg.num_experts_per_tok, routed_scaling_factor=1.0, topk_method="greedy",
                          norm_topk_prob=False, hidden_act="silu", n_group=1, topk_group=1)
    sd = {"model.embed_tokens.weight": ck["model.embed_tokens.weight"].float(), "model.norm.weight": ck["model.norm.weight"].float(),
          "lm_head.weight": ck["lm_head.weight"].float()}
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        for n in ("input_layernorm.weight", "post_attention_layernorm.weight", "self_attn.q_proj.weight", "self_attn.k_proj.weight",
                  "self_attn.v_proj.weight", "self_attn.o_proj.weight"):
            sd[p + n] = ck[p + n].float()
        if i < cfg.first_k_dense_replace:
            for n in ("gate_proj", "up_proj", "down_proj"):
                sd[p + f"mlp.{n}.weight"] = ck[p + f"mlp.{n}.weight"].float()
        else:
            moe = DeepseekV2Moe(dc).eval()
            msd = {"gate.weight": ck[p + "mlp.gate.weight"].float()}
            msd["experts.gate_up_proj"] = torch.stack([torch.cat([ck[p + f"mlp.experts.{e}.gate_proj.weight"].float(),
                                                                  ck[p + f"mlp.experts.{e}.up_proj.weight"].float()], 0)
                                                       for e in range(cfg.n_routed_experts)])
            msd["experts.down_proj"] = torch.stack([ck[p + f"mlp.experts.{e}.down_proj.weight"].float() for e in range(cfg.n_routed_experts)])
            for n in ("gate_proj", "up_proj", "down_proj"):
                msd[f"shared_experts.{n}.weight"] = ck[p + f"mlp.shared_experts.{n}.weight"].float()
            moe.load_state_dict(msd, strict=True)
            model.model.layers[i].mlp = moe
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(".mlp." in m for m in missing), (missing, unexpected)
    return model


def decoder_case(cfg, seed=5, n_img=20, steps=8):
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(2, cfg.vocab_size - 2, (9,), generator=g).tolist()
    ids, mask = D.build_prompt_tokens([[3, 4], text], [n_img], cfg)
    rows = torch.randn(n_img, cfg.hidden_size, generator=g) * 0.7
    forced = torch.randint(2, cfg.vocab_size - 2, (steps,), generator=g).tolist()
    return ids, mask, rows, forced


def decoder_reference_logits(model, ck, ids, mask, rows, forced):
    from transformers import DynamicCache

    emb = ck["model.embed_tokens.weight"].float()[torch.tensor(ids)]
    emb[torch.tensor(mask, dtype=torch.bool)] = rows
    cache = DynamicCache()
    out = model(inputs_embeds=emb.unsqueeze(0), past_key_values=cache, use_cache=True)
    logits = [out.logits[0, -1]]
    for t in forced[:-1]:
        e = ck["model.embed_tokens.weight"].float()[torch.tensor([[t]])]
        out = model(inputs_embeds=e, past_key_values=cache, use_cache=True)
        logits.append(out.logits[0, -1])
    return torch.stack(logits)


def checkpoint_digest(ck) -> str:
    h = hashlib.sha256()
    for k in ("model.sam_model.blocks.0.attn.qkv.weight", "model.vision_model.transformer.layers.1.mlp.fc1.weight",
              "model.layers.1.mlp.experts.3.down_proj.weight", "lm_head.weight"):
        h.update(ck[k].float().numpy().tobytes())
    return h.hexdigest()


def view(size, seed):
    page = P.synthetic_page(size + 13, size - 7, seed=seed)
    vi = P.prepare_vision_input(page, size, size, False)
    return torch.from_numpy(P.image_to_tensor(vi["global"])).unsqueeze(0)


def run(cfg, ck, store: bool):
    torch.manual_seed(0)
    ns = load_vllm_deepencoder()
    out = {}
    with torch.no_grad():
        sam = sam_reference(ns, cfg, ck)
        so, co = V.SamOracle(cfg, ck), V.ClipOracle(cfg, ck)
        for size in (640, 1024):
            x = view(size, seed=size)
            ref = sam(x)
            got = so.forward(x)
            cref = clip_reference(ns, cfg, ck, ref)
            cgot = co.forward(got)
            print(f"SAM  {size}: vLLM vs oracle max-abs {float((ref - got).abs().max()):.3e} (max |ref| {float(ref.abs().max()):.3f})")
            print(f"CLIP {size}: HF+vLLM vs oracle max-abs {float((cref - cgot).abs().max()):.3e} (max |ref| {float(cref.abs().max()):.3f})")
            out[f"sam_{size}"] = ref[0, ::4].numpy().copy()          # every 4th channel
            out[f"clip_{size}"] = cref[0, :, ::4].numpy().copy()
        model = decoder_reference(cfg, ck)
        ids, mask, rows, forced = decoder_case(cfg)
        ref = decoder_reference_logits(model, ck, ids, mask, rows, forced)
        lg = []
        D.DecoderOracle(cfg, ck).generate(ids, mask, rows, len(forced), 20, None, forced=forced, logits_out=lg)
        got = torch.stack(lg)
        print(f"DEC: HF Llama+DeepseekV2Moe vs oracle max-abs {float((ref - got).abs().max()):.3e} (max |ref| {float(ref.abs().max()):.3f}), "
              f"argmax equal {bool((ref.argmax(-1) == got.argmax(-1)).all())}")
        out["dec_logits"] = ref.numpy().copy()
    if store:
        out["digest"] = np.frombuffer(checkpoint_digest(ck).encode(), dtype=np.uint8)
        np.savez_compressed(HERE / "towers_tiny.npz", **out)
        print("wrote", HERE / "towers_tiny.npz", (HERE / "towers_tiny.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    cfg = OC.tiny_config()
    run(cfg, OC.random_checkpoint(cfg, seed=1234, storage=torch.bfloat16), store=True)
    if a.full:
        cfg = OC.full_config()
        cfg.vocab_size, cfg.image_token_id = 4096, 4095  # the lm_head / embedding width is not what this compares
        run(cfg, OC.random_checkpoint(cfg, seed=1234, storage=torch.bfloat16), store=False)
