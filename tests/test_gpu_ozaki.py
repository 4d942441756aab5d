"""tcgen05 Schur-complement kernel (russell_b200/csrc/ozaki_tc.cuh): int8 slices on tcgen05.mma.kind::i8, accumulators in
tensor memory, operands by TMA bulk copies, f64 recombination -- against numpy's f64 GEMM.

Tolerance: both operands are cut after 8 x 7 = 56 bits relative to their ROW maximum, and slice products with s + t > 9 are
dropped: |error(i, j)| <= ~8 * 2^-53 * K * max_k|a_ik| * max_k|b_jk| (the same shape as the rounding error of an f64 dot
product); the test allows 32 * 2^-53 * K * rowmax * colmax."""
import ctypes

import numpy as np
import pytest

import russell_b200 as rb
from russell_b200._lib import p_f64, ptr

pytestmark = pytest.mark.gpu


def _gemm(u, k, a, b, c):
    lib = rb._lib.load()
    ms = ctypes.c_double(0.0)
    af, bf, cf = np.asfortranarray(a), np.asfortranarray(b), np.asfortranarray(c)
    rc = lib.solver_b200_ozaki_gemm(u, k, ptr(af, p_f64), ptr(bf, p_f64), ptr(cf, p_f64), ctypes.byref(ms))
    assert rc == 0, rc
    return cf, ms.value


@pytest.mark.parametrize("u,k", [(128, 64), (64, 32), (300, 64), (129, 37), (1000, 128), (777, 200)])
def test_ozaki_gemm_matches_f64(u, k):
    rng = np.random.default_rng(u * 1000 + k)
    # rows of very different magnitude (the per-row exponents matter) and entries of very different magnitude inside a row
    a = rng.standard_normal((u, k)) * np.exp(rng.uniform(-20, 20, size=(u, 1))) * np.exp(rng.uniform(-6, 0, size=(u, k)))
    b = rng.standard_normal((u, k)) * np.exp(rng.uniform(-20, 20, size=(u, 1)))
    c = rng.standard_normal((u, u))
    got, _ = _gemm(u, k, a, b, c)
    want = c - a @ b.T
    bound = 32 * 2.0 ** -53 * k * np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :] + 4 * 2.0 ** -53 * np.abs(want)
    assert np.all(np.abs(got - want) <= bound), float(np.max(np.abs(got - want) / bound))


def test_ozaki_gemm_exact_on_small_integers_and_zero_rows():
    u, k = 200, 64
    rng = np.random.default_rng(5)
    a = rng.integers(-1000, 1000, size=(u, k)).astype(np.float64)
    b = rng.integers(-1000, 1000, size=(u, k)).astype(np.float64)
    a[17] = 0.0  # a zero row must not produce NaN (exponent of 0)
    b[3] = 0.0
    c = np.zeros((u, u))
    got, _ = _gemm(u, k, a, b, c)
    assert np.array_equal(got, -(a @ b.T))  # integers below 2^56 per row maximum are split without loss
