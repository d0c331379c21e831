"""Parity of the tcgen05 vision attention kernel (csrc/attention_tc.cuh) against the oracle's statement of
SAM attention with decomposed rel-pos bias (oracle/vision.py, following vision/sam.rs:804-888, 1124-1247) and
of CLIP attention (vision/clip.rs:349-381).  Tolerance: outputs are 16-bit and P is rounded to 16 bits before
P.V, so max-abs <= 3 * 2^-8 * max|out| for bf16 (2^-11 for f16), cosine >= 0.9999."""
import ctypes
import math

import numpy as np
import pytest
import torch

from dsocr.binding import lib, check
from oracle.vision import get_rel_pos

pytestmark = pytest.mark.gpu
BF16, F16 = 2, 1


def _round(t, dtype):
    return t.to(torch.bfloat16 if dtype == BF16 else torch.float16).to(torch.float32)


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def run_attn(dtype, qkv, B, S, H, grid, rel_h=None, rel_w=None):
    out = np.zeros((B * S, H * 64), dtype=np.float32)
    q = np.ascontiguousarray(qkv.numpy())
    rh = np.ascontiguousarray(rel_h.numpy()) if rel_h is not None else None
    rw = np.ascontiguousarray(rel_w.numpy()) if rel_w is not None else None
    st = lib().dsocr_test_vision_attention(dtype, B, S, H, _fp(q), grid, _fp(rh), _fp(rw),
                                           rel_h.shape[0] if rel_h is not None else 0, _fp(out))
    check(st, "dsocr_test_vision_attention")
    return torch.from_numpy(out)


def ref_attn(dtype, qkv, B, S, H, grid, rel_h=None, rel_w=None):
    x = _round(qkv, dtype).reshape(B, S, 3, H, 64).double()
    q, k, v = [x[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    s = (q @ k.transpose(2, 3)) * 0.125
    if grid:
        rh = get_rel_pos(grid, grid, _round(rel_h, dtype)).double()
        rw = get_rel_pos(grid, grid, _round(rel_w, dtype)).double()
        q5 = q.reshape(B, H, grid, grid, 64)
        bh = torch.einsum("bnhwd,hkd->bnhwk", q5, rh)
        bw = torch.einsum("bnhwd,wkd->bnhwk", q5, rw)
        s = s + (bh.unsqueeze(-1) + bw.unsqueeze(-2)).reshape(B, H, S, S)
    o = torch.softmax(s, -1) @ v
    return o.permute(0, 2, 1, 3).reshape(B * S, H * 64).float()


def _check(y, r, dtype):
    eps = 2 ** -8 if dtype == BF16 else 2 ** -11
    err = (y - r).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(y.flatten(), r.flatten(), dim=0).item()
    assert err <= 3 * eps * r.abs().max().item() + 1e-6, (err, r.abs().max().item())
    assert cos > 0.9999, cos


@pytest.mark.parametrize("dtype", [BF16, F16])
@pytest.mark.parametrize("S", [101, 257, 128, 40])
def test_clip_attention(dtype, S):
    g = torch.Generator().manual_seed(S)
    B, H = 2, 4
    qkv = torch.randn(B * S, 3 * H * 64, generator=g)
    _check(run_attn(dtype, qkv, B, S, H, 0), ref_attn(dtype, qkv, B, S, H, 0), dtype)


@pytest.mark.parametrize("dtype", [BF16, F16])
@pytest.mark.parametrize("grid", [14, 16, 32, 40, 64])
def test_sam_attention_relpos(dtype, grid):
    g = torch.Generator().manual_seed(grid)
    S = grid * grid
    B, H = (3, 2) if grid == 14 else (1, 2)
    qkv = torch.randn(B * S, 3 * H * 64, generator=g)
    rel_h = torch.randn(2 * grid - 1, 64, generator=g) * 0.2
    rel_w = torch.randn(2 * grid - 1, 64, generator=g) * 0.2
    y = run_attn(dtype, qkv, B, S, H, grid, rel_h, rel_w)
    r = ref_attn(dtype, qkv, B, S, H, grid, rel_h, rel_w)
    _check(y, r, dtype)


def test_sam_attention_peaked_scores():
    """Large-magnitude scores: exercises the online-softmax rescale across key blocks."""
    g = torch.Generator().manual_seed(1)
    grid, B, H = 32, 1, 1
    S = grid * grid
    qkv = torch.randn(B * S, 3 * H * 64, generator=g)
    qkv[:, :128] *= 4.0
    rel_h = torch.randn(2 * grid - 1, 64, generator=g)
    rel_w = torch.randn(2 * grid - 1, 64, generator=g)
    y = run_attn(BF16, qkv, B, S, H, grid, rel_h, rel_w)
    r = ref_attn(BF16, qkv, B, S, H, grid, rel_h, rel_w)
    _check(y, r, BF16)


@pytest.mark.parametrize("grid", [0, 40, 64])
def test_attention_lazy_rescale_non_uniform(grid):
    """Key norms ramp up along the sequence and only some query rows are large, so the reference maximum of
    individual rows (not whole warps) moves by more than the lazy-rescale threshold at many key blocks."""
    g = torch.Generator().manual_seed(7 + grid)
    S = grid * grid if grid else 777
    B, H = 1, 2
    qkv = torch.randn(B * S, 3, H, 64, generator=g)
    ramp = torch.linspace(0.2, 5.0, S).reshape(S, 1, 1)
    qkv[:, 1] *= ramp                       # keys grow along the sequence
    qkv[::3, 0] *= 6.0                      # every third query row is "hot"
    qkv = qkv.reshape(B * S, 3 * H * 64)
    rel_h = torch.randn(2 * grid - 1, 64, generator=g) * 0.3 if grid else None
    rel_w = torch.randn(2 * grid - 1, 64, generator=g) * 0.3 if grid else None
    y = run_attn(BF16, qkv, B, S, H, grid, rel_h, rel_w)
    r = ref_attn(BF16, qkv, B, S, H, grid, rel_h, rel_w)
    _check(y, r, BF16)
