"""CPU simulation of the GEMM precision modes of oryon_b200/csrc/gemm.cuh on one CLIP-sized linear layer (torch float8 dtypes,
float64 accumulation): relative RMS error of  one fp16 product  /  fp16 hi*hi + both cross terms in one 8-bit product (precision 2:
activations e5m2 with fixed scales, weights e4m3 behind a per-tensor power of two)  /  three fp16 products,  for activations of
different magnitude, one outlier channel and one outlier weight.  Run:  python tools/f8_cross_sim.py"""
import math

import torch


def f16(x):
    return x.to(torch.float32).to(torch.float16).to(torch.float64)


def e4m3(x):
    return x.to(torch.float32).clamp(-448, 448).to(torch.float8_e4m3fn).to(torch.float32).to(torch.float64)


def e5m2(x):
    return x.to(torch.float32).clamp(-57344, 57344).to(torch.float8_e5m2).to(torch.float32).to(torch.float64)


def main():
    torch.manual_seed(0)
    M, K, N = 512, 1024, 1024
    W = torch.randn(N, K, dtype=torch.float64) * 0.02
    W[3, 7] = 1.5
    print("activation scale | one product | fp8 cross terms | three products")
    for xs in (1e-2, 1.0, 1e2):
        X = torch.randn(M, K, dtype=torch.float64) * torch.exp(torch.randn(1, K, dtype=torch.float64) * 0.7) * xs
        X[:, 5] *= 30
        Y = X @ W.T
        rel = lambda a: ((a - Y).pow(2).mean().sqrt() / Y.pow(2).mean().sqrt()).item()
        k = 15 - math.frexp(W.abs().max().item())[1]          # gemm::weight_scale
        Wp, s = W * 2.0 ** k, 2.0 ** -k
        Xh, Wh = f16(X), f16(Wp)
        Xl, Wl = X - Xh, Wp - Wh
        p1 = rel((Xh @ Wh.T) * s)
        p3 = rel((Xh @ Wh.T + f16(Xl) @ Wh.T + Xh @ f16(Wl).T) * s)
        a_hi8, a_lo8 = e5m2(X * 2.0 ** -4), e5m2(Xl * 2.0 ** 7)      # kF8ActHi, kF8ActLo
        w_lo8, w_hi8 = e4m3(Wl * 2.0 ** 4), e4m3(Wp * 2.0 ** -7)     # kF8WLo, kF8WHi
        p2 = rel((Xh @ Wh.T + a_hi8 @ w_lo8.T + a_lo8 @ w_hi8.T) * s)
        print(f"{xs:16g} | {p1:11.2e} | {p2:15.2e} | {p3:14.2e}")


if __name__ == "__main__":
    main()
