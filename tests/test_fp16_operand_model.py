"""CPU model of the packed-FP16 operand arithmetic of the tensor-core kNN (csrc/knn_tc.cu: f16_scale,
split_f16_kernel, norm_pieces, pack_xyz_f16_kernel), in numpy float16 / float32 with the kernels'
rounding (round-to-nearest-even conversions, fp32 subtraction).  Pins the claims DESIGN.md §4 makes about
them without a GPU; the kernels themselves are checked against the oracle by the -m gpu tests.
"""
import math

import numpy as np
import pytest


def f16_scale(amax: float) -> float:
    """2^(12 - e) with amax = m * 2^e, m in [0.5, 1): the largest magnitude lands in [2^11, 2^12)."""
    if not (amax > 0.0) or not math.isfinite(amax):
        return 1.0
    _, e = math.frexp(amax)
    return math.ldexp(1.0, min(12 - e, 100))


def split(x: np.ndarray, s: float):
    v = (np.float32(s) * x.astype(np.float32)).astype(np.float32)
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    return v, hi, lo


@pytest.mark.parametrize("scale", [1.0, 1e-6, 3e4, 1e-20, 1e18])
def test_scaled_pair_keeps_22_bits_or_the_subnormal_floor(scale):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(20000) * np.exp(rng.uniform(-12, 0, 20000)) * scale).astype(np.float32)
    s = f16_scale(float(np.abs(x).max()))
    v, hi, lo = split(x, s)
    top = np.abs(v).max()
    assert 2.0 ** 11 <= top < 2.0 ** 12                                   # never near fp16's overflow (65504)
    assert np.isfinite(hi.astype(np.float32)).all() and np.isfinite(lo.astype(np.float32)).all()
    err = np.abs(v.astype(np.float64) - (hi.astype(np.float64) + lo.astype(np.float64)))
    bound = np.maximum(2.0 ** -22 * np.abs(v.astype(np.float64)), 2.0 ** -25)
    assert (err <= bound * 1.0000001).all(), float((err / bound).max())
    # scaling by a power of two commutes with fp32 rounding: the scaled squared norm is exactly s^2 |x|^2
    q = np.float32(0)
    for t in v[:64]:
        q = np.float32(t * t + q)
    q_ref = np.float32(0)
    for t in x[:64]:
        q_ref = np.float32(np.float32(t) * np.float32(t) + q_ref)
    if 1e-30 < float(q_ref) < 1e30:        # (unless the UNscaled squares leave fp32's normal range: the scaled ones never do)
        assert float(q) == float(q_ref) * s * s


@pytest.mark.parametrize("C", [64, 128])
def test_three_term_product_is_as_accurate_as_3xtf32(C):
    rng = np.random.default_rng(1)
    a = np.maximum(rng.standard_normal((64, C)), 0.2 * rng.standard_normal((64, C))).astype(np.float32)   # post-LeakyReLU-like
    s = f16_scale(float(np.abs(a).max()))
    v, hi, lo = split(a, s)
    H, L = hi.astype(np.float64), lo.astype(np.float64)
    three = H @ H.T + H @ L.T + L @ H.T                                   # what the three MMAs accumulate
    exact = (v.astype(np.float64) @ v.astype(np.float64).T)
    norm2 = (v.astype(np.float64) ** 2).sum(1)
    err = np.abs(three - exact).max() / norm2.max()
    assert err < 2e-6, err                                                # DESIGN: <= 1.2e-6 measured on the GPU
    # single-term first sweep: |hi.hi - exact| within the margin the kernel subtracts from its threshold
    one = H @ H.T
    nv = np.sqrt(norm2)
    margin = 1.1 * 2.0 ** -10 * nv[:, None] * nv.max() + 2.0 ** -24 * math.sqrt(2 * (C // 2)) * (nv[:, None] + nv.max())
    assert (np.abs(one - exact) <= margin).all()


@pytest.mark.parametrize("C", [64, 128])
def test_folded_column_term_fits_fp16_and_is_exact_to_2_pow_minus_33(C):
    rng = np.random.default_rng(2)
    top = np.float32(2.0 ** 12 * (1 - 2.0 ** -12))
    rows = [np.full(C, top, np.float32),                                   # the worst case: every channel at the range's top
            (rng.standard_normal(C) * 800).astype(np.float32), np.zeros(C, np.float32),
            np.full(C, 2.0 ** -20, np.float32)]
    for v in rows:
        q = np.float32((v.astype(np.float64) ** 2).sum())                 # |s x_j|^2 <= 2^31 for C <= 128
        assert q <= 2.0 ** 31
        w = np.float32(-q * np.float32(1.0 / 65536.0))
        p = []
        for _ in range(3):
            h = np.float16(w)
            p.append(h)
            w = np.float32(w - np.float32(h))
        assert all(np.isfinite(np.float32(t)) for t in p)                 # |q| / 2^16 <= 2^15 < 65504
        got = 2.0 ** 15 * sum(float(t) for t in p)                        # 2^15 on the query side, pieces on the other
        want = -0.5 * float(q)
        assert abs(got - want) <= max(2.0 ** -33 * abs(want), 2.0 ** -25 * 2.0 ** 15), (got, want)


def test_xyz_rows_put_the_whole_score_into_one_16_deep_k_step():
    rng = np.random.default_rng(3)
    pts = rng.standard_normal((50, 3)).astype(np.float32)
    pts /= np.linalg.norm(pts, axis=1).max()
    s = f16_scale(float(np.abs(pts).max()))
    v, hi, lo = split(pts, s)
    C = 3
    A = np.zeros((50, 16), np.float64)
    B = np.zeros((50, 16), np.float64)
    for n in range(50):
        q = np.float32(0)
        for c in range(C):
            q = np.float32(v[n, c] * v[n, c] + q)
        w = np.float32(-q * np.float32(1.0 / 65536.0))
        pc = []
        for _ in range(3):
            h = np.float16(w)
            pc.append(float(h))
            w = np.float32(w - np.float32(h))
        A[n, 0:3], A[n, 3:6], A[n, 6:9], A[n, 9:12] = hi[n], hi[n], lo[n], 2.0 ** 15
        B[n, 0:3], B[n, 3:6], B[n, 6:9], B[n, 9:12] = hi[n], lo[n], hi[n], pc
    score = A @ B.T                                                       # one K = 16 MMA per tile
    vd = v.astype(np.float64)
    exact = vd @ vd.T - 0.5 * (vd ** 2).sum(1)[None, :]
    assert np.abs(score - exact).max() <= 2e-6 * (vd ** 2).sum(1).max()
    # the ranking the kernel extracts equals the exact one (self first, same neighbour sets for k = 8)
    k = 8
    mine = np.argsort(-score, axis=1, kind="stable")[:, :k]
    ref = np.argsort(-exact, axis=1, kind="stable")[:, :k]
    assert (np.sort(mine, 1) == np.sort(ref, 1)).mean() > 0.99
    assert (mine[:, 0] == np.arange(50)).all()


def test_degenerate_scales():
    assert f16_scale(0.0) == 1.0 and f16_scale(float("inf")) == 1.0 and f16_scale(float("nan")) == 1.0
    assert f16_scale(1e-44) == 2.0 ** 100                                 # a tensor of denormals keeps a finite scale
    for a in (1.0, 0.75, 3.9999, 4.0, 1e-7, 6e4):
        assert 2.0 ** 11 <= a * f16_scale(a) < 2.0 ** 12
