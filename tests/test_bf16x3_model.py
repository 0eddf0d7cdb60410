"""CPU: the arithmetic model of the tcgen05 likelihood (turing.jl_b200/csrc/logistic_tc.cuh) restated in
numpy -- operands split into three bf16 planes h + m + l, the six plane products down to 2^-16 * 2^-8
accumulated in fp32, for D > 128 the two feature halves accumulated separately and added (what the CTA
pair does) -- against the fp64 dot product.  It pins the bound the header states for eta,
|d eta| <= 1e-6 * sum_k |theta_k| |x_k|, independently of any GPU, and shows what the dropped products
would cost (the reason GEMM1 keeps six of the nine)."""
import numpy as np
import pytest


def bf16(x):
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def split3(v):
    v = np.asarray(v, dtype=np.float64)
    h = bf16(v)
    m = bf16(v - h)
    l = bf16(v - h - m.astype(np.float64))
    return h, m, l


def eta_bf16x3(theta, X, products, halves=1):
    """fp32 accumulation of the listed (theta plane, x plane) products; `halves` feature blocks are
    accumulated separately and added in fp32 (one CTA per block)"""
    tp, xp = split3(theta), split3(X)
    D = theta.shape[-1]
    edges = np.linspace(0, D, halves + 1).astype(int) if halves > 1 else np.array([0, D])
    if halves == 2:
        edges = np.array([0, min(128, D), D])
    total = np.zeros(X.shape[0], dtype=np.float32)
    for a, b in zip(edges[:-1], edges[1:]):
        acc = np.zeros(X.shape[0], dtype=np.float32)
        for (pa, pb) in products:                       # smallest products first, as the kernel issues them
            for k in range(a, b):
                acc = (acc + (xp[pb][:, k] * tp[pa][k]).astype(np.float32)).astype(np.float32)
        total = (total + acc).astype(np.float32)
    return total


SIX = [(2, 0), (1, 1), (0, 2), (1, 0), (0, 1), (0, 0)]     # l*h, m*m, h*l, m*h, h*m, h*h
THREE = [(1, 0), (0, 1), (0, 0)]


@pytest.mark.parametrize("D,halves", [(100, 1), (128, 1), (200, 2), (256, 2)])
def test_six_plane_products_are_an_fp32_grade_dot_product(D, halves):
    rng = np.random.default_rng(D)
    X = rng.standard_normal((300, D))
    theta = rng.standard_normal(D) * 0.5
    ref = X @ theta
    scale = np.abs(X) @ np.abs(theta)
    err6 = np.abs(eta_bf16x3(theta, X, SIX, halves).astype(np.float64) - ref) / scale
    assert err6.max() < 1e-6, err6.max()                  # the bound stated in logistic_tc.cuh
    # the planes are exact: h + m + l reproduces every operand to 2^-24 relative
    h, m, l = split3(X)
    rec = h.astype(np.float64) + m + l
    assert np.max(np.abs(rec - X) / np.abs(X)) < 2.0 ** -23
    # dropping the 2^-16 products (what GEMM2 does for the gradient, where it is harmless) is visible here
    err3 = np.abs(eta_bf16x3(theta, X, THREE, halves).astype(np.float64) - ref) / scale
    assert err3.max() > 2 * err6.max() and err3.max() < 3e-5
