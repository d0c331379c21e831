"""f32 CPU restatement of the vision path (SAM ViT-B -> CLIP-L -> projector -> token layout).

Test infrastructure only.  Follows, op for op:
  * crates/infer-deepseek/src/vision/sam.rs   (whole file)
  * crates/infer-deepseek/src/vision/clip.rs  (whole file)
  * crates/infer-deepseek/src/model/mod.rs:392-444, 590-923 (projector + token formatting)
Candle semantics restated: layer_norm (biased variance, eps inside sqrt), ops::softmax
(max-subtracted), gelu_erf (exact), conv2d (cross-correlation, zero pad).
Liberty taken (kinder than the reference, numerically equivalent): the decomposed rel-pos bias
is a vectorised einsum instead of the scalar host loop at sam.rs:1148-1186.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .config import OcrConfig


def _f32(ckpt, name):
    return ckpt[name].to(torch.float32)


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def bicubic_resize_antialiased(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """sam.rs:1000-1123, written out (f32, Pillow filter a=-0.5, vertical pass then horizontal)."""
    def filt(t: torch.Tensor) -> torch.Tensor:
        a = -0.5
        t = t.abs()
        r1 = ((a + 2.0) * t - (a + 3.0)) * t * t + 1.0
        r2 = (((t - 5.0) * t + 8.0) * t - 4.0) * a
        return torch.where(t < 1.0, r1, torch.where(t < 2.0, r2, torch.zeros_like(t)))

    def axis_matrix(in_len: int, out_len: int) -> torch.Tensor:
        scale = torch.tensor(in_len / out_len, dtype=torch.float32)
        support = 2.0 * scale if scale >= 1.0 else torch.tensor(2.0)
        invscale = 1.0 / scale if scale >= 1.0 else torch.tensor(1.0)
        m = torch.zeros(out_len, in_len, dtype=torch.float32)
        for o in range(out_len):
            center = scale * (o + 0.5)
            xmin = max(int(math.floor(float(center - support + 0.5))), 0)
            xmax = min(int(math.floor(float(center + support + 0.5))), in_len)
            n = max(xmax - xmin, 0)
            if n == 0:
                continue
            j = torch.arange(n, dtype=torch.float32)
            w = filt((j + (xmin - center) + 0.5) * invscale)
            tot = w.sum()
            if tot != 0:
                w = w / tot
            m[o, xmin:xmax] = w
        return m

    _, _, in_h, in_w = x.shape
    if in_h == out_h and in_w == out_w:
        return x
    my = axis_matrix(in_h, out_h)
    mx = axis_matrix(in_w, out_w)
    tmp = torch.einsum("oh,bchw->bcow", my, x.to(torch.float32))
    return torch.einsum("pw,bcow->bcop", mx, tmp)


def get_rel_pos(q_size: int, k_size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """sam.rs:1194-1247 -> [q_size, k_size, head_dim] (linear interp, align_corners=False)."""
    orig_len, hd = rel_pos.shape
    max_rel = 2 * max(q_size, k_size) - 1
    rel = rel_pos.to(torch.float32)
    if orig_len != max_rel:
        scale = torch.tensor(orig_len / max_rel, dtype=torch.float32)
        i = torch.arange(max_rel, dtype=torch.float32)
        src = (scale * (i + 0.5) - 0.5).clamp(0.0, float(orig_len - 1))
        left = src.floor()
        li = left.long()
        ri = (li + 1).clamp(max=orig_len - 1)
        w = (src - left).clamp(0.0, 1.0).unsqueeze(1)
        rel = rel[li] * (1.0 - w) + rel[ri] * w
    scale_q = max(k_size / q_size, 1.0)
    scale_k = max(q_size / k_size, 1.0)
    qi = torch.arange(q_size, dtype=torch.float32).unsqueeze(1) * scale_q
    ki = torch.arange(k_size, dtype=torch.float32).unsqueeze(0) * scale_k
    idx = ((qi - ki) + (k_size - 1.0) * scale_k).floor().clamp(0, max_rel - 1).long()
    return rel[idx]


def window_partition(x: torch.Tensor, window: int):
    """sam.rs:926-955."""
    b, h, w, c = x.shape
    pad_h = (window - h % window) % window
    pad_w = (window - w % window) % window
    if pad_h or pad_w:
        x = F.pad(x, (0, 0, 0, pad_w, 0, pad_h))
    hp, wp = h + pad_h, w + pad_w
    x = x.reshape(b, hp // window, window, wp // window, window, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, window, window, c), (hp, wp)


def window_unpartition(windows: torch.Tensor, window: int, pad_hw, hw):
    """sam.rs:957-980."""
    hp, wp = pad_hw
    h, w = hw
    c = windows.shape[-1]
    b = windows.shape[0] // ((hp // window) * (wp // window))
    x = windows.reshape(b, hp // window, wp // window, window, window, c).permute(0, 1, 3, 2, 4, 5)
    x = x.reshape(b, hp, wp, c)
    return x[:, :h, :w, :].contiguous()


class SamOracle:
    def __init__(self, cfg: OcrConfig, ckpt: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.p = {k[len("model.sam_model."):]: v.to(torch.float32) for k, v in ckpt.items()
                  if k.startswith("model.sam_model.")}

    def attention(self, i: int, x: torch.Tensor, spatial) -> torch.Tensor:
        """sam.rs:804-888 + compute_relative_bias :1124-1192.  x: [B, H, W, C]."""
        cfg, p = self.cfg, self.p
        b, h, w, c = x.shape
        nh = cfg.sam_heads
        hd = c // nh
        s = h * w
        pre = f"blocks.{i}.attn."
        qkv = F.linear(x.reshape(b, s, c), p[pre + "qkv.weight"], p[pre + "qkv.bias"])
        qkv = qkv.reshape(b, s, 3, nh, hd)
        q = qkv[:, :, 0].permute(0, 2, 1, 3)
        k = qkv[:, :, 1].permute(0, 2, 1, 3)
        v = qkv[:, :, 2].permute(0, 2, 1, 3)
        scores = (q @ k.transpose(2, 3)) * (1.0 / math.sqrt(hd))
        qh, qw = spatial
        rh = get_rel_pos(qh, qh, p[pre + "rel_pos_h"])  # [qh, kh, hd]
        rw = get_rel_pos(qw, qw, p[pre + "rel_pos_w"])
        q5 = q.reshape(b, nh, qh, qw, hd)
        rel_h = torch.einsum("bnhwd,hkd->bnhwk", q5, rh)
        rel_w = torch.einsum("bnhwd,wkd->bnhwk", q5, rw)
        bias = (rel_h.unsqueeze(-1) + rel_w.unsqueeze(-2)).reshape(b, nh, s, s)
        attn = torch.softmax(scores + bias, dim=-1)
        ctx = (attn @ v).permute(0, 2, 1, 3).reshape(b, h, w, c)
        return F.linear(ctx, p[pre + "proj.weight"], p[pre + "proj.bias"])

    def block(self, i: int, x: torch.Tensor) -> torch.Tensor:
        """sam.rs:731-748."""
        cfg, p = self.cfg, self.p
        pre = f"blocks.{i}."
        b, h, w, c = x.shape
        normed = layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-6)
        if i in cfg.sam_global_idx:
            a = self.attention(i, normed, (h, w))
        else:
            win, pad_hw = window_partition(normed, cfg.sam_window)
            a = self.attention(i, win, (cfg.sam_window, cfg.sam_window))
            a = window_unpartition(a, cfg.sam_window, pad_hw, (h, w))
        x = x + a
        n2 = layer_norm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-6)
        m = F.linear(n2, p[pre + "mlp.lin1.weight"], p[pre + "mlp.lin1.bias"])
        m = F.gelu(m)  # exact erf (sam.rs:921)
        m = F.linear(m, p[pre + "mlp.lin2.weight"], p[pre + "mlp.lin2.bias"])
        return x + m

    def forward(self, img: torch.Tensor, trace: Optional[dict] = None) -> torch.Tensor:
        """sam.rs:210-289.  img [B,3,H,W] f32 -> [B, 1024, H/64, W/64]."""
        cfg, p = self.cfg, self.p
        x = F.conv2d(img, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=cfg.sam_patch)
        x = x.permute(0, 2, 3, 1).contiguous()
        if trace is not None:
            trace["patch_embed"] = x.clone()
        _, th, tw, _ = x.shape
        pos = p["pos_embed"]
        if pos.shape[1] != th or pos.shape[2] != tw:
            pos = bicubic_resize_antialiased(pos.permute(0, 3, 1, 2), th, tw).permute(0, 2, 3, 1)
        x = x + pos
        if trace is not None:
            trace["pos_added"] = x.clone()
            trace["block_outputs"] = []
        for i in range(cfg.sam_depth):
            x = self.block(i, x)
            if trace is not None:
                trace["block_outputs"].append(x.clone())
        x = x.permute(0, 3, 1, 2).contiguous()

        def ln2d(t, w, b):
            return layer_norm(t.permute(0, 2, 3, 1), w, b, 1e-6).permute(0, 3, 1, 2)

        c1 = F.conv2d(x, p["neck.0.weight"])
        n1 = ln2d(c1, p["neck.1.weight"], p["neck.1.bias"])
        c2 = F.conv2d(n1, p["neck.2.weight"], padding=1)
        n2 = ln2d(c2, p["neck.3.weight"], p["neck.3.bias"])
        d2 = F.conv2d(n2, p["net_2.weight"], stride=2, padding=1)
        d3 = F.conv2d(d2, p["net_3.weight"], stride=2, padding=1)
        if trace is not None:
            trace.update(neck_conv1=c1, neck_norm1=n1, neck_conv2=c2, neck_norm2=n2, net2=d2, net3=d3)
        return d3


class ClipOracle:
    def __init__(self, cfg: OcrConfig, ckpt: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.p = {k[len("model.vision_model."):]: v.to(torch.float32) for k, v in ckpt.items()
                  if k.startswith("model.vision_model.")}

    def adapt_pos(self, target_tokens: int) -> torch.Tensor:
        """clip.rs:486-544."""
        table = self.p["embeddings.position_embedding.weight"]
        src_tokens, hidden = table.shape
        if src_tokens == target_tokens:
            return table
        src = int(round(math.sqrt(src_tokens - 1)))
        tgt = int(round(math.sqrt(target_tokens - 1)))
        grid = table[1:].reshape(src, src, hidden).permute(2, 0, 1).unsqueeze(0)
        res = bicubic_resize_antialiased(grid, tgt, tgt).squeeze(0).permute(1, 2, 0).reshape(tgt * tgt, hidden)
        return torch.cat([table[:1], res], 0)

    def forward(self, sam_out: torch.Tensor, trace: Optional[dict] = None) -> torch.Tensor:
        """clip.rs:98-102, 165-236.  sam_out [B, 1024, g, g] -> [B, g*g+1, 1024] (no final LN)."""
        cfg, p = self.cfg, self.p
        b, c, gh, gw = sam_out.shape
        patches = sam_out.reshape(b, c, gh * gw).transpose(1, 2)
        cls = p["embeddings.class_embedding"].reshape(1, 1, c).expand(b, 1, c)
        x = torch.cat([cls, patches], 1) + self.adapt_pos(gh * gw + 1).unsqueeze(0)
        if trace is not None:
            trace["embeddings"] = x.clone()
        x = layer_norm(x, p["pre_layrnorm.weight"], p["pre_layrnorm.bias"], 1e-5)
        if trace is not None:
            trace["pre_layernorm"] = x.clone()
            trace["layer_outputs"] = []
        nh = cfg.clip_heads
        hd = c // nh
        for i in range(cfg.clip_layers):
            pre = f"transformer.layers.{i}."
            n = layer_norm(x, p[pre + "layer_norm1.weight"], p[pre + "layer_norm1.bias"], 1e-5)
            qkv = F.linear(n, p[pre + "self_attn.qkv_proj.weight"], p[pre + "self_attn.qkv_proj.bias"])
            s = qkv.shape[1]
            q, k, v = [t.reshape(b, s, nh, hd).permute(0, 2, 1, 3) for t in qkv.split(c, dim=-1)]
            attn = torch.softmax((q @ k.transpose(2, 3)) * (1.0 / math.sqrt(hd)), dim=-1)
            ctx = (attn @ v).transpose(1, 2).reshape(b, s, c)
            x = x + F.linear(ctx, p[pre + "self_attn.out_proj.weight"], p[pre + "self_attn.out_proj.bias"])
            n = layer_norm(x, p[pre + "layer_norm2.weight"], p[pre + "layer_norm2.bias"], 1e-5)
            m = F.linear(n, p[pre + "mlp.fc1.weight"], p[pre + "mlp.fc1.bias"])
            m = m * torch.sigmoid(1.702 * m)  # quick-GELU clip.rs:413-416
            x = x + F.linear(m, p[pre + "mlp.fc2.weight"], p[pre + "mlp.fc2.bias"])
            if trace is not None:
                trace["layer_outputs"].append(x.clone())
        return x


class VisionOracle:
    """model/mod.rs:526-924 (VisionContext) for OCR-1."""

    def __init__(self, cfg: OcrConfig, ckpt: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.sam = SamOracle(cfg, ckpt)
        self.clip = ClipOracle(cfg, ckpt)
        self.proj_w = _f32(ckpt, "model.projector.layers.weight")
        self.proj_b = _f32(ckpt, "model.projector.layers.bias")
        self.newline = _f32(ckpt, "model.image_newline")
        self.separator = _f32(ckpt, "model.view_seperator")

    def pre_tokens(self, img: torch.Tensor) -> torch.Tensor:
        """build_clip_sam_tokens model/mod.rs:604-650: [CLIP(1024) | SAM(1024)] per token."""
        sam = self.sam.forward(img)
        clip = self.clip.forward(sam)
        b, c, gh, gw = sam.shape
        return torch.cat([clip[:, 1:], sam.reshape(b, c, gh * gw).transpose(1, 2)], -1)

    def project(self, pre: torch.Tensor) -> torch.Tensor:
        return F.linear(pre, self.proj_w, self.proj_b)

    def _row_breaks(self, grid: torch.Tensor) -> torch.Tensor:
        rows, cols, hid = grid.shape
        nl = self.newline.reshape(1, 1, hid).expand(rows, 1, hid)
        return torch.cat([grid, nl], 1).reshape(rows * (cols + 1), hid)

    def encode(self, global_chw: torch.Tensor, patches: Optional[torch.Tensor], crop_shape, taps=None
               ) -> torch.Tensor:
        """compute_image_embeddings for one page -> [N_img, 1280] = [local ; global ; view_separator]."""
        if global_chw.dim() == 3:
            global_chw = global_chw.unsqueeze(0)
        gpre = self.pre_tokens(global_chw)
        gpost = self.project(gpre)
        side = int(math.isqrt(gpost.shape[1]))
        gtok = self._row_breaks(gpost[0].reshape(side, side, -1))
        segs = []
        if patches is not None and patches.shape[0] > 0:
            wc, hc = crop_shape
            lpre = self.pre_tokens(patches)
            lpost = self.project(lpre)
            s = int(math.isqrt(lpost.shape[1]))
            grid = lpost.reshape(hc, wc, s, s, -1).permute(0, 2, 1, 3, 4).reshape(hc * s, wc * s, -1)
            segs.append(self._row_breaks(grid))
            if taps is not None:
                taps.update(local_pre=lpre, local_post=lpost)
        segs.append(gtok)
        segs.append(self.separator.reshape(1, -1))
        if taps is not None:
            taps.update(global_pre=gpre, global_post=gpost)
        return torch.cat(segs, 0)
