"""GPU parity of the backbone's tensor-core GEMM (oryon_gemm_f32) against a float64 torch evaluation of
``residual + act(alpha * A W^T + bias)``.

Tolerances, relative to the magnitude scale ``sqrt(K) * rms(A) * rms(W)`` of one output:
  precision 3 (fp16 split pairs, three tcgen05 products)  1e-4   operand error is ~2^-21; what remains is the tensor
                                                                  core's truncating fp32 accumulation, ~(K/16) * 2^-24 per
                                                                  output (measured 2e-5 at K = 1024)
  precision 1 (single fp16 product)                        2e-3
  precision 2 (one fp16 product + both cross terms in      2e-4   cross terms carry ~2^-15 relative error per element: 1/18 of the
               one 8-bit product, K % 64 == 0)                     single-product error (tools/f8_cross_sim.py), whatever the
                                                                  magnitude of the activations (e5m2 needs no scale)
"""
import pytest
import torch

from gpu_util import need_gpu
from oryon_b200 import ops

pytestmark = pytest.mark.gpu


def _ref(A, W, bias, res, act, alpha):
    y = alpha * (A.double() @ W.double().transpose(-1, -2))
    if bias is not None:
        y = y + bias.double()
    if act == "quickgelu":
        y = y * torch.sigmoid(1.702 * y)
    elif act == "gelu":
        y = torch.nn.functional.gelu(y)
    elif act == "relu":
        y = torch.relu(y)
    if res is not None:
        y = y + res.double()
    return y


CASES = [
    # M, N, K, batch, act, bias, residual
    (256, 128, 64, 1, "none", False, False),
    (577, 1024, 1024, 1, "none", True, True),
    (1154, 3072, 1024, 1, "none", True, False),
    (300, 4096, 1024, 1, "quickgelu", True, False),
    (300, 1000, 4096, 1, "none", True, True),
    (576, 80, 768, 3, "none", False, False),      # cost volume: per-image text matrix, N < tile
    (2304, 64, 1152, 1, "relu", True, False),     # decoder conv as GEMM, TN = 64
    (9216, 16, 1152, 1, "relu", True, False),     # TN = 32, N = 16
    (100, 33, 48, 2, "gelu", True, True),         # everything ragged
    (1, 7, 5, 1, "none", True, False),
    # large, N % 256 == 0: CTA pairs (gemm_tc2_kernel, tcgen05 cta_group::2; >= 74 tiles of 256 x 256); with ORYON_GEMM_QUAD=1 and
    # N % 512 == 0: two CTA pairs per cluster sharing A through TMA multicast (gemm_tc4_kernel), see test_quad_kernel_... below
    (4800, 1024, 1024, 1, "none", True, True),        # out-projection shape, ragged last row block (4800 = 18.75 x 256)
    (4737, 3072, 1024, 1, "none", True, False),       # QKV shape, one row in the last block
    (5000, 4096, 1024, 1, "quickgelu", True, False),  # MLP up-projection
    (4864, 1024, 4096, 1, "none", True, True),        # MLP down-projection, deep K
    (2500, 512, 320, 2, "gelu", True, True),          # batched, K not a multiple of 64
    (19000, 256, 64, 1, "relu", False, False),        # one k-block, one column tile (pair kernel)
    (10000, 768, 256, 1, "none", True, True),         # N % 512 != 0: pair kernel, three column tiles
    (9700, 512, 128, 1, "gelu", True, False),         # exactly one unit per row block (quad kernel), ragged rows
]


@pytest.mark.parametrize("M,N,K,batch,act,use_bias,use_res", CASES)
@pytest.mark.parametrize("precision", [3, 1])
def test_gemm_matches_float64(M, N, K, batch, act, use_bias, use_res, precision):
    need_gpu()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + batch)
    shape_a = (batch, M, K) if batch > 1 else (M, K)
    shape_w = (batch, N, K) if batch > 1 else (N, K)
    A = torch.randn(shape_a, generator=g)
    W = torch.randn(shape_w, generator=g) * 0.05
    bias = torch.randn(N, generator=g) if use_bias else None
    res = torch.randn(*shape_a[:-1], N, generator=g) if use_res else None
    alpha = 0.125 if act == "none" and not use_bias else 1.0
    out = ops.linear(A.cuda(), W.cuda(), None if bias is None else bias.cuda(), None if res is None else res.cuda(), act=act,
                     alpha=alpha, precision=precision).cpu()
    ref = _ref(A, W, bias, res, act, alpha)
    scale = alpha * (K ** 0.5) * 1.0 * 0.05 + 1e-3
    err = (out.double() - ref).abs().max().item() / scale
    tol = 1e-4 if precision == 3 else 2e-3
    print(f'scaled err {err:.3e}')
    assert err < tol, f"max scaled error {err:.3e} (tol {tol})"


@pytest.mark.parametrize("M,N,K,batch,act,use_bias,use_res", [c for c in CASES if c[2] % 64 == 0])
def test_gemm_fp8_cross_terms_match_float64(M, N, K, batch, act, use_bias, use_res):
    """precision 2 on every case whose depth is a multiple of 64: both kernels (single CTA for the small cases, CTA pairs for the
    large ones), batched operands, every epilogue."""
    need_gpu()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + batch)
    shape_a = (batch, M, K) if batch > 1 else (M, K)
    shape_w = (batch, N, K) if batch > 1 else (N, K)
    A = torch.randn(shape_a, generator=g)
    W = torch.randn(shape_w, generator=g) * 0.05
    bias = torch.randn(N, generator=g) if use_bias else None
    res = torch.randn(*shape_a[:-1], N, generator=g) if use_res else None
    alpha = 0.125 if act == "none" and not use_bias else 1.0
    out = ops.linear(A.cuda(), W.cuda(), None if bias is None else bias.cuda(), None if res is None else res.cuda(), act=act,
                     alpha=alpha, precision=2).cpu()
    ref = _ref(A, W, bias, res, act, alpha)
    scale = alpha * (K ** 0.5) * 1.0 * 0.05 + 1e-3
    err = (out.double() - ref).abs().max().item() / scale
    print(f'scaled err {err:.3e}')
    assert err < 2e-4, f"max scaled error {err:.3e} (tol 2e-4)"


@pytest.mark.parametrize("act_scale", [1e-2, 1.0, 1e2])
def test_gemm_fp8_cross_terms_any_activation_scale(act_scale):
    """The activation side of the 8-bit product is e5m2 with fixed scales, the weight side e4m3 behind a per-tensor power of two taken
    from max|W|: the error relative to the output scale must not depend on the magnitude of the activations, nor suffer from one
    outlier channel (30x) or one outlier weight (75x the rms)."""
    need_gpu()
    g = torch.Generator().manual_seed(11)
    M, N, K = 4864, 1024, 1024
    A = torch.randn(M, K, generator=g) * torch.exp(0.7 * torch.randn(1, K, generator=g)) * act_scale
    A[:, 5] *= 30.0
    W = torch.randn(N, K, generator=g) * 0.02
    W[3, 7] = 1.5
    ref = A.double() @ W.double().T
    out2 = ops.linear(A.cuda(), W.cuda(), precision=2).cpu().double()
    out1 = ops.linear(A.cuda(), W.cuda(), precision=1).cpu().double()
    out3 = ops.linear(A.cuda(), W.cuda(), precision=3).cpu().double()
    rms = ref.pow(2).mean().sqrt()
    e1, e2, e3 = [((o - ref).pow(2).mean().sqrt() / rms).item() for o in (out1, out2, out3)]
    print(f"relative rms error: one product {e1:.2e}, fp8 cross terms {e2:.2e}, three products {e3:.2e}")
    assert e2 < 4e-5 and e2 < e1 / 8, (e1, e2, e3)


def test_gemm_fp8_cross_terms_need_a_depth_multiple_of_64():
    """The 8-bit cross-term blocks cover whole 64-deep K blocks: other depths are refused, not padded silently."""
    need_gpu()
    from oryon_b200._lib import OryonError
    A, W = torch.randn(64, 100).cuda(), torch.randn(32, 100).cuda()
    with pytest.raises(OryonError):
        ops.linear(A, W, precision=2)
    assert torch.isfinite(ops.linear(A, W, precision=3)).all()     # the handle stays usable


def test_gemm_large_values_saturate_not_nan():
    need_gpu()
    A = torch.full((128, 64), 300.0)
    W = torch.full((128, 64), 300.0)
    out = ops.linear(A.cuda(), W.cuda()).cpu()
    assert torch.isfinite(out).all() and torch.allclose(out, torch.full_like(out, 300.0 * 300.0 * 64), rtol=1e-6)


@pytest.mark.parametrize("M,N,K,batch,act,use_bias,use_res", [c for c in CASES if c[1] % 512 == 0 and c[0] >= 4000])
def test_quad_kernel_matches_float64(M, N, K, batch, act, use_bias, use_res, monkeypatch):
    """gemm_tc4_kernel (two CTA pairs per cluster, A through TMA multicast; opt-in because it is slower on the B200's 148 SMs) is
    held to the same tolerance as the default kernels."""
    monkeypatch.setenv("ORYON_GEMM_QUAD", "1")
    test_gemm_matches_float64(M, N, K, batch, act, use_bias, use_res, 3)
    test_gemm_matches_float64(M, N, K, batch, act, use_bias, use_res, 1)
