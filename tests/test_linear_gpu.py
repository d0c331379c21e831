"""Parity of the tcgen05 linear kernel (csrc/linear_tc.cuh) against an f32 torch reference of the same op
on operands rounded to the tensor-core type.  Tolerances: f32 outputs 2e-3 relative to the row scale
(accumulation-order only); 16-bit outputs additionally one rounding (2^-8 bf16 / 2^-11 f16)."""
import ctypes

import numpy as np
import pytest
import torch

from dsocr.binding import lib, check

pytestmark = pytest.mark.gpu

BF16, F16 = 2, 1


def _round(t, dtype):
    return t.to(torch.bfloat16 if dtype == BF16 else torch.float16).to(torch.float32)


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def run_linear(dtype, x, w0, w1=None, bias=None, act=0, out_mode=2, x_parts=1, bn=0, row_map=None, out_init=None,
               out_rows=None):
    M, K = x.shape
    N = w0.shape[0]
    out_rows = out_rows or M
    out = np.zeros((out_rows, N), dtype=np.float32) if out_init is None else out_init.copy()
    xs = np.ascontiguousarray(x.numpy())
    w0s = np.ascontiguousarray(w0.numpy())
    w1s = np.ascontiguousarray(w1.numpy()) if w1 is not None else None
    bs = np.ascontiguousarray(bias.numpy()) if bias is not None else None
    rm = np.ascontiguousarray(row_map.astype(np.int32)) if row_map is not None else None
    st = lib().dsocr_test_linear(dtype, M, N, K, _fp(xs), _fp(w0s), _fp(w1s), _fp(bs), act, out_mode, x_parts, bn,
                                 rm.ctypes.data_as(ctypes.POINTER(ctypes.c_int)) if rm is not None else None,
                                 out_rows, _fp(out))
    check(st, "dsocr_test_linear")
    return torch.from_numpy(out)


def ref_linear(dtype, x, w0, w1=None, bias=None, act=0, x_parts=1):
    xr = x if x_parts == 2 else _round(x, dtype)
    if x_parts == 2:
        hi = _round(x, dtype)
        xr = hi + _round(x - hi, dtype)
    y = xr.double() @ _round(w0, dtype).double().T
    if bias is not None:
        y = y + bias.double()
    if w1 is not None:
        u = xr.double() @ _round(w1, dtype).double().T
        y = torch.nn.functional.silu(y) * u
    elif act == 1:
        y = torch.nn.functional.gelu(y)
    elif act == 2:
        y = y * torch.sigmoid(1.702 * y)
    return y.float()


SHAPES = [(1, 128, 64), (7, 256, 128), (33, 384, 192), (64, 1280, 1280), (130, 200, 320), (300, 2304, 768),
          (1000, 768, 3072), (257, 6848, 1280)]


@pytest.mark.parametrize("dtype", [BF16, F16])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_f32_out(dtype, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.05
    b = torch.randn(N, generator=g)
    y = run_linear(dtype, x, w, bias=b)
    r = ref_linear(dtype, x, w, bias=b)
    assert torch.allclose(y, r, rtol=1e-3, atol=2e-3 * r.abs().max().item()), (y - r).abs().max()


@pytest.mark.parametrize("bn", [32, 64, 128, 256])
def test_linear_all_token_tiles(bn):
    g = torch.Generator().manual_seed(bn)
    x = torch.randn(517, 256, generator=g)
    w = torch.randn(384, 256, generator=g) * 0.05
    y = run_linear(BF16, x, w, bn=bn)
    r = ref_linear(BF16, x, w)
    assert torch.allclose(y, r, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("act", [1, 2])
def test_linear_16bit_out_activations(act):
    g = torch.Generator().manual_seed(act)
    x = torch.randn(400, 768, generator=g)
    w = torch.randn(512, 768, generator=g) * 0.04
    b = torch.randn(512, generator=g) * 0.1
    y = run_linear(BF16, x, w, bias=b, act=act, out_mode=0)
    r = ref_linear(BF16, x, w, bias=b, act=act)
    assert torch.allclose(y, r, rtol=2 ** -7, atol=1e-2)


@pytest.mark.parametrize("dtype", [BF16, F16])
def test_linear_split_activations_near_f32(dtype):
    """hi/lo split activations: result matches the f32-activation product to ~2^-16 relative."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(70, 1280, generator=g)
    w = torch.randn(640, 1280, generator=g) * 0.02
    y = run_linear(dtype, x, w, x_parts=2)
    exact = (x.double() @ _round(w, dtype).double().T).float()
    plain = ref_linear(dtype, x, w)
    err_split = (y - exact).abs().max().item()
    err_plain = (plain - exact).abs().max().item()
    assert err_split < 2e-4 * exact.abs().max().item()
    assert err_split < err_plain / 20


@pytest.mark.parametrize("x_parts", [1, 2])
@pytest.mark.parametrize("M", [5, 64, 200])
def test_linear_swiglu_dual(M, x_parts):
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, 1280, generator=g)
    w0 = torch.randn(896, 1280, generator=g) * 0.03
    w1 = torch.randn(896, 1280, generator=g) * 0.03
    y = run_linear(BF16, x, w0, w1=w1, out_mode=1, x_parts=x_parts)
    r = ref_linear(BF16, x, w0, w1=w1, x_parts=x_parts)
    assert torch.allclose(y, r, rtol=1e-3, atol=2e-3 * r.abs().max().item())


def test_linear_residual_add_with_row_map():
    g = torch.Generator().manual_seed(11)
    M, N, K = 392, 768, 768
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.03
    b = torch.randn(N, generator=g)
    rm = np.full(M, -1, dtype=np.int32)
    keep = np.random.RandomState(0).permutation(M)[:300]
    rm[keep] = np.arange(300)
    base = torch.randn(300, N, generator=g)
    y = run_linear(BF16, x, w, bias=b, out_mode=3, row_map=rm, out_init=base.numpy(), out_rows=300)
    r = base.clone()
    r[torch.from_numpy(rm[keep]).long()] += ref_linear(BF16, x, w, bias=b)[torch.from_numpy(keep).long()]
    assert torch.allclose(y, r, rtol=1e-3, atol=5e-3)


@pytest.mark.parametrize("dual", [False, True])
def test_grouped_linear_matches_per_expert(dual):
    g = torch.Generator().manual_seed(3)
    E, N, K = 16, 896 if dual else 1280, 1280 if dual else 896
    counts = np.array([0, 1, 7, 33, 64, 65, 0, 130, 2, 3, 4, 5, 0, 0, 300, 9], dtype=np.int32)
    M = int(counts.sum())
    x = torch.randn(M, K, generator=g)
    w0 = torch.randn(E, N, K, generator=g) * 0.03
    w1 = torch.randn(E, N, K, generator=g) * 0.03 if dual else None
    out = np.zeros((M, N), dtype=np.float32)
    st = lib().dsocr_test_grouped_linear(
        BF16, E, M, N, K, counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _fp(np.ascontiguousarray(x.numpy())),
        _fp(np.ascontiguousarray(w0.numpy())), _fp(np.ascontiguousarray(w1.numpy())) if dual else None, 2, _fp(out))
    check(st, "dsocr_test_grouped_linear")
    y = torch.from_numpy(out)
    r = torch.zeros(M, N)
    row = 0
    for e in range(E):
        c = int(counts[e])
        if c:
            r[row:row + c] = ref_linear(BF16, x[row:row + c], w0[e], w1=w1[e] if dual else None, x_parts=2)
        row += c
    assert torch.allclose(y, r, rtol=1e-3, atol=2e-3 * r.abs().max().item())


@pytest.mark.parametrize("dual,cap,sparse", [(False, 64, False), (True, 64, False), (True, 24, False), (False, 200, False),
                                             (True, 130, False), (True, 64, True), (False, 8, True)])
def test_fixed_capacity_grouped_linear_device_scheduled(dual, cap, sparse):
    """Decode-time expert GEMM: units are enumerated on the device from the per-expert counts (empty experts and
    partially filled segments in any mix) and their k-blocks are dealt out evenly to the CTAs (stream-K with a
    fixed-order fix-up); result must equal the per-expert reference, untouched rows stay zero.  (The reuse of
    the hand-off flags across launches is covered by the decoder tests, which replay the step hundreds of times.)"""
    g = torch.Generator().manual_seed(11 + cap)
    E, N, K = 64, 896 if dual else 1280, 1280 if dual else 896
    counts = torch.randint(0, min(cap, 12) + 1, (E,), generator=g).numpy().astype(np.int32)
    counts[[0, 5, 6, 7, 63]] = 0
    counts[9] = cap
    counts[40] = max(1, cap - 1)
    if sparse:  # a handful of populated experts: every unit's reduction is split over several CTAs (stream-K)
        keep = counts[[9, 21, 40]].copy()
        counts[:] = 0
        counts[[9, 21, 40]] = np.maximum(keep, 1)
    x = torch.randn(E * cap, K, generator=g)
    w0 = torch.randn(E, N, K, generator=g) * 0.03
    w1 = torch.randn(E, N, K, generator=g) * 0.03 if dual else None
    out = np.zeros((E * cap, N), dtype=np.float32)
    st = lib().dsocr_test_fixedcap_linear(
        BF16, E, cap, N, K, counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _fp(np.ascontiguousarray(x.numpy())),
        _fp(np.ascontiguousarray(w0.numpy())), _fp(np.ascontiguousarray(w1.numpy())) if dual else None, 2, _fp(out))
    check(st, "dsocr_test_fixedcap_linear")
    y = torch.from_numpy(out)
    r = torch.zeros(E * cap, N)
    for e in range(E):
        c = int(counts[e])
        if c:
            r[e * cap:e * cap + c] = ref_linear(BF16, x[e * cap:e * cap + c], w0[e], w1=w1[e] if dual else None, x_parts=2)
    assert torch.allclose(y, r, rtol=1e-3, atol=2e-3 * r.abs().max().item())
    for e in range(E):
        assert not y[e * cap + int(counts[e]):(e + 1) * cap].any()


# ---- CTA-pair kernel (csrc/linear_pair.cuh): picked for M >= 2048 tokens and N % 256 == 0 ----------------------
@pytest.mark.parametrize("dtype", [BF16, F16])
@pytest.mark.parametrize("M,N,K", [(2048, 256, 64), (4173, 768, 768), (2304, 2304, 768), (3001, 1024, 4096)])
def test_linear_pair_f32_out(dtype, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.05
    b = torch.randn(N, generator=g)
    y = run_linear(dtype, x, w, bias=b)
    r = ref_linear(dtype, x, w, bias=b)
    assert torch.allclose(y, r, rtol=1e-3, atol=2e-3 * r.abs().max().item()), (y - r).abs().max()


@pytest.mark.parametrize("act", [1, 2])
def test_linear_pair_16bit_out_activations(act):
    g = torch.Generator().manual_seed(40 + act)
    x = torch.randn(2500, 768, generator=g)
    w = torch.randn(3072, 768, generator=g) * 0.04
    b = torch.randn(3072, generator=g) * 0.1
    y = run_linear(BF16, x, w, bias=b, act=act, out_mode=0)
    r = ref_linear(BF16, x, w, bias=b, act=act)
    assert torch.allclose(y, r, rtol=2 ** -7, atol=1e-2)


@pytest.mark.parametrize("M,N,K", [(2500, 768, 3072), (4096, 1024, 1024), (2049, 768, 768)])
def test_linear_pair_residual_add_plain(M, N, K):
    """x += y through the bulk-reduction epilogue (no row map): row tails of the last 16-token chunk and of the last
    256-token tile, every element updated exactly once."""
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.03
    b = torch.randn(N, generator=g)
    base = torch.randn(M, N, generator=g)
    y = run_linear(BF16, x, w, bias=b, out_mode=3, out_init=base.numpy())
    r = base + ref_linear(BF16, x, w, bias=b)
    assert torch.allclose(y, r, rtol=1e-3, atol=5e-3), (y - r).abs().max()


def test_linear_pair_residual_add_with_row_map():
    g = torch.Generator().manual_seed(12)
    M, N, K = 2940, 768, 768   # 15 windows of 14x14 tokens, some of them padding
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.03
    b = torch.randn(N, generator=g)
    rm = np.full(M, -1, dtype=np.int32)
    keep = np.random.RandomState(1).permutation(M)[:2500]
    rm[keep] = np.arange(2500)
    base = torch.randn(2500, N, generator=g)
    y = run_linear(BF16, x, w, bias=b, out_mode=3, row_map=rm, out_init=base.numpy(), out_rows=2500)
    r = base.clone()
    r[torch.from_numpy(rm[keep]).long()] += ref_linear(BF16, x, w, bias=b)[torch.from_numpy(keep).long()]
    assert torch.allclose(y, r, rtol=1e-3, atol=5e-3)
