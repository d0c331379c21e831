"""Dequant-fused tensor-core GEMM (csrc/linear_dq.cuh; run_quantized_matmul, quantization.rs:164-185) against
y = (x_hi + x_lo) . dequant(W)^T with the oracle's dequantisers (pinned to gguf-py in tests/test_golden_cpu.py), per
format, for the shapes the decoder uses: single and dual (SwiGLU) weights, every token tile, split-K, table-grouped
experts, N and M that are not multiples of the tiles.  Tolerance: 2e-4 of the output scale (weights are split into
hi + lo 16-bit parts, accumulation is f32; measured ~1e-5 bf16 / ~1e-6 f16)."""
import ctypes

import numpy as np
import pytest
import torch

from dsocr.binding import check, lib
from oracle import dsq

pytestmark = pytest.mark.gpu
BF16, F16 = 2, 1
QUANT = {dsq.Q8_0: dsq.quantize_q8_0, dsq.Q4K: dsq.quantize_q4k, dsq.Q6K: dsq.quantize_q6k}


def _round_split(x, dtype):
    t = torch.bfloat16 if dtype == BF16 else torch.float16
    hi = x.to(t).float()
    return hi + (x - hi).to(t).float()


def _run(dtype, qd, x, blocks, blocks1, N, groups=1, counts=None, bn=0, k_splits=1):
    M, K = x.shape
    out = np.zeros((M, N), np.float32)
    xs = np.ascontiguousarray(x.numpy())
    u8 = ctypes.POINTER(ctypes.c_uint8)
    b0 = np.frombuffer(blocks, np.uint8)
    b1 = np.frombuffer(blocks1, np.uint8) if blocks1 is not None else None
    cn = (ctypes.c_int * groups)(*counts) if counts is not None else None
    check(lib().dsocr_test_linear_dq(dtype, qd, groups, cn, M, N, K, b0.ctypes.data_as(u8),
                                     b1.ctypes.data_as(u8) if b1 is not None else None,
                                     xs.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), bn, k_splits,
                                     out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))), "dsocr_test_linear_dq")
    return torch.from_numpy(out)


def _weights(qd, rows, K, seed):
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(rows, K, generator=g) * 0.05).numpy()
    blocks = QUANT[qd](w)
    return blocks, torch.from_numpy(dsq.dequantize(blocks, qd, rows, K))


@pytest.mark.parametrize("dtype", [BF16, F16])
@pytest.mark.parametrize("qd", [dsq.Q8_0, dsq.Q4K, dsq.Q6K])
@pytest.mark.parametrize("M,N,K,bn", [(300, 1280, 1280, 0), (40, 384, 1792, 32), (130, 200, 512, 64), (700, 256, 768, 256)])
def test_single_weight(dtype, qd, M, N, K, bn):
    blocks, wd = _weights(qd, N, K, seed=M + N)
    x = torch.randn(M, K, generator=torch.Generator().manual_seed(K))
    got = _run(dtype, qd, x, blocks, None, N, bn=bn)
    ref = (_round_split(x, dtype).double() @ wd.double().T).float()
    err = (got - ref).abs().max().item()
    print(f"[parity] linear_dq q{qd} dtype {dtype} {M}x{N}x{K} bn {bn}: max-abs {err:.3e} (scale {ref.abs().max().item():.3f})")
    assert err <= 2e-4 * ref.abs().max().item()


@pytest.mark.parametrize("qd", [dsq.Q8_0, dsq.Q4K, dsq.Q6K])
def test_dual_swiglu_and_split_k(qd):
    M, N, K = 24, 896, 1280
    b0, w0 = _weights(qd, N, K, seed=1)
    b1, w1 = _weights(qd, N, K, seed=2)
    x = torch.randn(M, K, generator=torch.Generator().manual_seed(3))
    xr = _round_split(x, BF16).double()
    ref = (torch.nn.functional.silu(xr @ w0.double().T) * (xr @ w1.double().T)).float()
    for ks in (1, 4):
        got = _run(BF16, qd, x, b0, b1, N, bn=32, k_splits=ks)
        err = (got - ref).abs().max().item()
        print(f"[parity] linear_dq dual q{qd} k_splits {ks}: max-abs {err:.3e} (scale {ref.abs().max().item():.3f})")
        assert err <= 2e-4 * ref.abs().max().item()
    got = _run(BF16, qd, x, b0, None, N, bn=32, k_splits=5)
    ref1 = (xr @ w0.double().T).float()
    assert (got - ref1).abs().max().item() <= 2e-4 * ref1.abs().max().item()


@pytest.mark.parametrize("qd,K,N", [(dsq.Q4K, 1280, 896), (dsq.Q8_0, 896, 1280)])
def test_grouped_experts(qd, K, N):
    """Table-grouped form of the MoE prefill: stacked expert matrices, ragged row counts incl. empty experts."""
    counts = [5, 0, 131, 64, 1, 0, 17, 300]
    groups, M = len(counts), sum(counts)
    blocks, wd = _weights(qd, groups * N, K, seed=9)
    x = torch.randn(M, K, generator=torch.Generator().manual_seed(4))
    got = _run(F16, qd, x, blocks, None, N, groups=groups, counts=counts, bn=128)
    xr = _round_split(x, F16).double()
    ref = torch.zeros(M, N)
    r = 0
    for g, c in enumerate(counts):
        ref[r:r + c] = (xr[r:r + c] @ wd[g * N:(g + 1) * N].double().T).float()
        r += c
    err = (got - ref).abs().max().item()
    print(f"[parity] linear_dq grouped q{qd}: max-abs {err:.3e} (scale {ref.abs().max().item():.3f})")
    assert err <= 2e-4 * ref.abs().max().item()
