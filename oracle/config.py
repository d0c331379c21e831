"""Model configuration + random-init checkpoint generator.  Test infrastructure only.

Architecture constants: SURVEY.md section 8 "Architecture constants" (pinned where the
reference tests pin them: crates/infer-deepseek/tests/config.rs:32-56,
tests/vision_sam.rs:30-35, tests/vision_clip.rs:16-20).  Tensor names follow the
reference loaders: vision/sam.rs:143-184,483-500,545-546,762-801,897-915;
vision/clip.rs:73-88,130-163,281-299,329-346,391-403; model/mod.rs:272-307;
transformer/weights.rs:187-191,279-281,363-400,451-454,504-530.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field, asdict
from typing import Dict, List

import numpy as np
import torch


@dataclass
class OcrConfig:
    # SAM ViT-B (vision/sam.rs:44-111)
    sam_image_size: int = 1024
    sam_patch: int = 16
    sam_dim: int = 768
    sam_depth: int = 12
    sam_heads: int = 12
    sam_window: int = 14
    sam_global_idx: List[int] = field(default_factory=lambda: [2, 5, 8, 11])
    sam_neck: int = 256
    sam_out: List[int] = field(default_factory=lambda: [512, 1024])
    # CLIP-L (vision/clip.rs:35-52)
    clip_dim: int = 1024
    clip_layers: int = 24
    clip_heads: int = 16
    clip_image_size: int = 224
    clip_patch: int = 14
    # projector (model/mod.rs:263-307)
    proj_in: int = 2048
    n_embed: int = 1280
    # decoder (transformer/weights.rs, config/mod.rs)
    vocab_size: int = 129280
    hidden_size: int = 1280
    num_layers: int = 12
    num_heads: int = 10
    intermediate_size: int = 6848
    moe_intermediate_size: int = 896
    n_routed_experts: int = 64
    n_shared_experts: int = 2
    num_experts_per_tok: int = 6
    first_k_dense_replace: int = 1
    rope_theta: float = 10000.0
    rms_norm_eps: float = 1e-6
    bos_token_id: int = 0
    eos_token_id: int = 1
    image_token_id: int = 128815

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    def to_reference_json(self) -> dict:
        """config.json in the schema config/mod.rs:38-66 parses."""
        return {
            "architectures": ["DeepseekOCRForCausalLM"],
            "model_type": "deepseek_vl_v2",
            "language_config": {
                "vocab_size": self.vocab_size,
                "hidden_size": self.hidden_size,
                "intermediate_size": self.intermediate_size,
                "moe_intermediate_size": self.moe_intermediate_size,
                "num_hidden_layers": self.num_layers,
                "num_attention_heads": self.num_heads,
                "num_key_value_heads": self.num_heads,
                "n_routed_experts": self.n_routed_experts,
                "n_shared_experts": self.n_shared_experts,
                "num_experts_per_tok": self.num_experts_per_tok,
                "first_k_dense_replace": self.first_k_dense_replace,
                "moe_layer_freq": 1,
                "norm_topk_prob": False,
                "routed_scaling_factor": 1.0,
                "scoring_func": "softmax",
                "topk_method": "greedy",
                "hidden_act": "silu",
                "rope_theta": self.rope_theta,
                "rms_norm_eps": self.rms_norm_eps,
                "q_lora_rank": None,
                "kv_lora_rank": None,
                "qk_rope_head_dim": 0,
                "qk_nope_head_dim": 0,
                "v_head_dim": 0,
                "use_mla": False,
                "max_position_embeddings": 8192,
                "bos_token_id": self.bos_token_id,
                "eos_token_id": self.eos_token_id,
                "torch_dtype": "bfloat16",
            },
            "projector_config": {"input_dim": self.proj_in, "n_embed": self.n_embed, "projector_type": "linear"},
            "vision_config": {
                "image_size": self.sam_image_size,
                "model_name": "deepencoder",
                "width": {
                    "sam_vit_b": {
                        "width": self.sam_dim, "layers": self.sam_depth, "heads": self.sam_heads,
                        "patch_size": self.sam_patch, "image_size": self.sam_image_size,
                        "global_attn_indexes": self.sam_global_idx, "downsample_channels": self.sam_out,
                    },
                    "clip-l-14-224": {
                        "width": self.clip_dim, "layers": self.clip_layers, "heads": self.clip_heads,
                        "patch_size": self.clip_patch, "image_size": self.clip_image_size,
                    },
                },
            },
            "image_token_id": self.image_token_id,
        }

    def save_json(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump(self.to_reference_json(), f, indent=1)


def full_config() -> OcrConfig:
    return OcrConfig()


def tiny_config(**kw) -> OcrConfig:
    """Same per-head / per-channel dimensions as the real model (the kernels are specialised on
    those) but few layers, few experts and a small vocabulary so the CPU oracle runs in seconds."""
    cfg = OcrConfig(
        sam_depth=3, sam_global_idx=[2], clip_layers=2, vocab_size=2048, num_layers=3,
        n_routed_experts=16, image_token_id=2047,
    )
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def _randn(gen: torch.Generator, shape, std: float) -> torch.Tensor:
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std


def tensor_shapes(cfg: OcrConfig) -> Dict[str, tuple]:
    """name -> (shape, kind) with kind in {w, b, norm_w, norm_b, table, emb, head}."""
    out: Dict[str, tuple] = {}
    s = "model.sam_model."
    g = cfg.sam_image_size // cfg.sam_patch
    out[s + "patch_embed.proj.weight"] = ((cfg.sam_dim, 3, cfg.sam_patch, cfg.sam_patch), "w")
    out[s + "patch_embed.proj.bias"] = ((cfg.sam_dim,), "b")
    out[s + "pos_embed"] = ((1, g, g, cfg.sam_dim), "table")
    for i in range(cfg.sam_depth):
        p = f"{s}blocks.{i}."
        rel = 2 * (g if i in cfg.sam_global_idx else cfg.sam_window) - 1
        hd = cfg.sam_dim // cfg.sam_heads
        for n in ("norm1", "norm2"):
            out[p + n + ".weight"] = ((cfg.sam_dim,), "norm_w")
            out[p + n + ".bias"] = ((cfg.sam_dim,), "norm_b")
        out[p + "attn.qkv.weight"] = ((3 * cfg.sam_dim, cfg.sam_dim), "w")
        out[p + "attn.qkv.bias"] = ((3 * cfg.sam_dim,), "b")
        out[p + "attn.proj.weight"] = ((cfg.sam_dim, cfg.sam_dim), "w")
        out[p + "attn.proj.bias"] = ((cfg.sam_dim,), "b")
        out[p + "attn.rel_pos_h"] = ((rel, hd), "table")
        out[p + "attn.rel_pos_w"] = ((rel, hd), "table")
        out[p + "mlp.lin1.weight"] = ((4 * cfg.sam_dim, cfg.sam_dim), "w")
        out[p + "mlp.lin1.bias"] = ((4 * cfg.sam_dim,), "b")
        out[p + "mlp.lin2.weight"] = ((cfg.sam_dim, 4 * cfg.sam_dim), "w")
        out[p + "mlp.lin2.bias"] = ((cfg.sam_dim,), "b")
    out[s + "neck.0.weight"] = ((cfg.sam_neck, cfg.sam_dim, 1, 1), "w")
    out[s + "neck.1.weight"] = ((cfg.sam_neck,), "norm_w")
    out[s + "neck.1.bias"] = ((cfg.sam_neck,), "norm_b")
    out[s + "neck.2.weight"] = ((cfg.sam_neck, cfg.sam_neck, 3, 3), "w")
    out[s + "neck.3.weight"] = ((cfg.sam_neck,), "norm_w")
    out[s + "neck.3.bias"] = ((cfg.sam_neck,), "norm_b")
    out[s + "net_2.weight"] = ((cfg.sam_out[0], cfg.sam_neck, 3, 3), "w")
    out[s + "net_3.weight"] = ((cfg.sam_out[1], cfg.sam_out[0], 3, 3), "w")

    c = "model.vision_model."
    npos = (cfg.clip_image_size // cfg.clip_patch) ** 2 + 1
    out[c + "embeddings.class_embedding"] = ((cfg.clip_dim,), "table")
    out[c + "embeddings.position_embedding.weight"] = ((npos, cfg.clip_dim), "table")
    out[c + "pre_layrnorm.weight"] = ((cfg.clip_dim,), "norm_w")
    out[c + "pre_layrnorm.bias"] = ((cfg.clip_dim,), "norm_b")
    for i in range(cfg.clip_layers):
        p = f"{c}transformer.layers.{i}."
        for n in ("layer_norm1", "layer_norm2"):
            out[p + n + ".weight"] = ((cfg.clip_dim,), "norm_w")
            out[p + n + ".bias"] = ((cfg.clip_dim,), "norm_b")
        out[p + "self_attn.qkv_proj.weight"] = ((3 * cfg.clip_dim, cfg.clip_dim), "w")
        out[p + "self_attn.qkv_proj.bias"] = ((3 * cfg.clip_dim,), "b")
        out[p + "self_attn.out_proj.weight"] = ((cfg.clip_dim, cfg.clip_dim), "w")
        out[p + "self_attn.out_proj.bias"] = ((cfg.clip_dim,), "b")
        out[p + "mlp.fc1.weight"] = ((4 * cfg.clip_dim, cfg.clip_dim), "w")
        out[p + "mlp.fc1.bias"] = ((4 * cfg.clip_dim,), "b")
        out[p + "mlp.fc2.weight"] = ((cfg.clip_dim, 4 * cfg.clip_dim), "w")
        out[p + "mlp.fc2.bias"] = ((cfg.clip_dim,), "b")

    out["model.projector.layers.weight"] = ((cfg.n_embed, cfg.proj_in), "w")
    out["model.projector.layers.bias"] = ((cfg.n_embed,), "b")
    out["model.image_newline"] = ((cfg.n_embed,), "emb")
    out["model.view_seperator"] = ((cfg.n_embed,), "emb")

    out["model.embed_tokens.weight"] = ((cfg.vocab_size, cfg.hidden_size), "emb")
    H = cfg.hidden_size
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        out[p + "input_layernorm.weight"] = ((H,), "norm_w")
        out[p + "post_attention_layernorm.weight"] = ((H,), "norm_w")
        for n in "qkvo":
            out[p + f"self_attn.{n}_proj.weight"] = ((H, H), "w")
        if i < cfg.first_k_dense_replace:
            I = cfg.intermediate_size
            out[p + "mlp.gate_proj.weight"] = ((I, H), "w")
            out[p + "mlp.up_proj.weight"] = ((I, H), "w")
            out[p + "mlp.down_proj.weight"] = ((H, I), "w")
        else:
            I = cfg.moe_intermediate_size
            out[p + "mlp.gate.weight"] = ((cfg.n_routed_experts, H), "gate")
            for e in range(cfg.n_routed_experts):
                q = f"{p}mlp.experts.{e}."
                out[q + "gate_proj.weight"] = ((I, H), "w")
                out[q + "up_proj.weight"] = ((I, H), "w")
                out[q + "down_proj.weight"] = ((H, I), "w")
            S = I * cfg.n_shared_experts
            out[p + "mlp.shared_experts.gate_proj.weight"] = ((S, H), "w")
            out[p + "mlp.shared_experts.up_proj.weight"] = ((S, H), "w")
            out[p + "mlp.shared_experts.down_proj.weight"] = ((H, S), "w")
    out["model.norm.weight"] = ((H,), "norm_w")
    out["lm_head.weight"] = ((cfg.vocab_size, H), "head")
    return out


# std per tensor kind.  Linears N(0, 0.02) (initializer_range default, config/mod.rs:309-311).
# Biases / norm offsets are given small non-zero values so that parity tests exercise them
# (an all-zero bias hides indexing bugs).  The router gate and the embedding rows get a larger
# scale so that routing / logits are not near-uniform (SURVEY.md section 7 "hard parts").
_KIND_STD = {"w": 0.02, "b": 0.02, "norm_b": 0.02, "table": 0.02, "emb": 0.5, "head": 0.02, "gate": 0.08}


def random_checkpoint(cfg: OcrConfig, seed: int = 1234, storage: torch.dtype = torch.bfloat16
                      ) -> Dict[str, torch.Tensor]:
    """Random-init checkpoint of the exact architecture, values already rounded to `storage`
    (the reference keeps f16/bf16 *storage* and computes in f32: SURVEY.md 8a 'dtype semantics')."""
    gen = torch.Generator().manual_seed(seed)
    ckpt: Dict[str, torch.Tensor] = {}
    for name, (shape, kind) in tensor_shapes(cfg).items():
        if kind == "norm_w":
            t = 1.0 + _randn(gen, shape, 0.05)
        else:
            t = _randn(gen, shape, _KIND_STD[kind])
        ckpt[name] = t.to(storage).contiguous()
    return ckpt


def save_checkpoint(ckpt: Dict[str, torch.Tensor], path: str) -> None:
    from safetensors.torch import save_file

    save_file(ckpt, path)


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    from safetensors.torch import load_file

    return load_file(path)
