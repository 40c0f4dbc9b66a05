"""CPU check of the arithmetic behind GEMM precision 2 (oryon_b200/csrc/gemm.cuh): x*w ~ hi_x*hi_w (fp16) + [x 2^-4 | (x - hi_x) 2^7] . [(w' - hi_w) 2^4 |
w' 2^-7] (8-bit, one product), with the activation side in e5m2 at FIXED scales and the weight side in e4m3 behind a per-tensor power of
two (max|w'| in (2^14, 2^15]).  torch's float8 dtypes on the CPU stand in for the tensor core; products are exact there as they are in
hardware, the accumulation is float64 here (the hardware's is fp32).  The GPU counterpart is tests/test_gemm_gpu.py
(test_gemm_fp8_cross_terms_*), which measured 1.67e-5 where this model says 1.65e-5."""
import math

import pytest
import torch


def _f16(x):
    return x.to(torch.float32).to(torch.float16).to(torch.float64)


def _e4m3(x):
    return x.to(torch.float32).clamp(-448, 448).to(torch.float8_e4m3fn).to(torch.float32).to(torch.float64)


def _e5m2(x):
    return x.to(torch.float32).clamp(-57344, 57344).to(torch.float8_e5m2).to(torch.float32).to(torch.float64)


def weight_scale(absmax: float) -> float:
    """gemm::weight_scale: 2^k with absmax * 2^k in [2^14, 2^15)."""
    return 2.0 ** (15 - math.frexp(absmax)[1])


def products(X, W):
    """(one fp16 product, precision 2, three fp16 products) of X @ W.T in the library's operand formats."""
    sc = weight_scale(W.abs().max().item())
    Wp = W * sc
    Xh, Wh = _f16(X), _f16(Wp)
    Xl, Wl = X - Xh, Wp - Wh
    p1 = (Xh @ Wh.T) / sc
    p3 = (Xh @ Wh.T + _f16(Xl) @ Wh.T + Xh @ _f16(Wl).T) / sc
    a_first, a_second = _e5m2(X * 2.0 ** -4), _e5m2(Xl * 2.0 ** 7)        # kF8ActHi, kF8ActLo
    w_first, w_second = _e4m3(Wl * 2.0 ** 4), _e4m3(Wp * 2.0 ** -7)       # kF8WLo, kF8WHi
    p2 = (Xh @ Wh.T + a_first @ w_first.T + a_second @ w_second.T) / sc
    return p1, p2, p3


@pytest.mark.parametrize("act_scale", [1e-2, 1.0, 1e2])
def test_cross_terms_in_eight_bits_recover_an_order_of_magnitude(act_scale):
    g = torch.Generator().manual_seed(0)
    M, K, N = 256, 1024, 512
    X = torch.randn(M, K, generator=g, dtype=torch.float64) * torch.exp(0.7 * torch.randn(1, K, generator=g, dtype=torch.float64)) * act_scale
    X[:, 5] *= 30.0                                   # an outlier channel
    W = torch.randn(N, K, generator=g, dtype=torch.float64) * 0.02
    W[3, 7] = 1.5                                     # an outlier weight: 75x the rms, it sets the scale
    Y = X @ W.T
    rel = lambda a: ((a - Y).pow(2).mean().sqrt() / Y.pow(2).mean().sqrt()).item()
    e1, e2, e3 = (rel(p) for p in products(X, W))
    assert 2e-4 < e1 < 4e-4                           # one product: 2^-11 per operand
    assert e2 < 2.5e-5 and e2 < e1 / 12               # the 8-bit cross terms: ~2^-15 per element, whatever the activation scale
    assert e3 < 2e-6


def test_weight_scale_is_a_power_of_two_that_fills_the_fp16_range():
    for amax in (1e-6, 0.02, 1.0, 1.5, 300.0, 16384.0, 40000.0):
        sc = weight_scale(amax)
        assert math.frexp(sc)[0] == 0.5               # a power of two: scaling and unscaling are exact
        assert 2 ** 14 <= amax * sc < 2 ** 15         # hi = fp16(w') never overflows, w' 2^-7 <= 256 fits e4m3, (w' - hi) 2^4 <= 128 too


def test_block_scales_multiply_to_one():
    # first halves: x 2^-4 with (w' - hi) 2^4; second halves: (x - hi) 2^7 with w' 2^-7 -- both products carry no scale
    assert 2.0 ** -4 * 2.0 ** 4 == 1.0 and 2.0 ** 7 * 2.0 ** -7 == 1.0
