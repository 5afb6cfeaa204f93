"""GPU tests of the moment schedule of the Gram stage (csrc/gram_moment.cu): the Gram matrix of a monomial
library is a moment matrix, every distinct moment is contracted once (virtual row / column libraries, only the
block pairs the staircase of needed entries touches) and gathered into the packed [G | R | yy].  It must agree
with the oracle (Theta Theta', Theta Y', diag Y Y' of reference STLSQ.jl:90-91 on the materialised library) and
with the plain schedules, and with itself bit for bit."""
import os

import numpy as np
import pytest

import ddb200
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ddb200.Context(0)
    yield c
    c.close()


def _gram(ctx, moment):
    old = os.environ.get("DDB_GRAM_MOMENT")
    os.environ["DDB_GRAM_MOMENT"] = "1" if moment else "0"
    try:
        ctx.gram()
        return ctx.gram_download(), ctx.timings()
    finally:
        if old is None:
            del os.environ["DDB_GRAM_MOMENT"]
        else:
            os.environ["DDB_GRAM_MOMENT"] = old


def _check(ctx, ex, X, Y, tol=1e-13):
    ctx.set_library(ex)
    ctx.upload(X, Y)
    (G1, R1, y1), t1 = _gram(ctx, True)
    (G0, R0, y0), t0 = _gram(ctx, False)
    Th = oracle.evaluate_library(ex, X)
    Gr, Rr, yr = Th @ Th.T, Th @ Y.T, (Y * Y).sum(1)
    sc = np.sqrt(np.outer(np.diag(Gr), np.diag(Gr)))
    for G, R, yy in ((G1, R1, y1), (G0, R0, y0)):
        assert np.max(np.abs(G - Gr) / sc) < tol
        assert np.max(np.abs(R - Rr)) < tol * np.sqrt(np.max(np.diag(Gr)) * np.max(yr))
        np.testing.assert_allclose(yy, yr, rtol=tol)
        np.testing.assert_array_equal(G, G.T)
    return (G1, R1, y1), t1, t0


@pytest.mark.parametrize("m", [4096, 20_001, 65_536 + 17])
def test_moment_schedule_lorenz96_library(ctx, m):
    """n = 64, degree 2 (k = 2145, config C3): ragged last stage, more / fewer stages than sample ranges."""
    n = 64
    rng = np.random.default_rng(m)
    X = rng.standard_normal((n, m))
    Y = rng.standard_normal((n, m))
    (G, R, yy), t1, t0 = _check(ctx, oracle.polynomial_exponents(n, 2), X, Y)
    assert t1["gram_launches"] == 2  # self-building kernel over the staircase + gather
    assert G[0, 0] == m  # constant x constant: a sum of ones is exact


def _plan_says_used(ex, n_t):
    """The host-only analysis (ddb_moment_plan_describe) plus the run-time size gate (more than two blocks)."""
    from ddb200 import _lib

    exf = np.asfortranarray(ex.astype(np.uint8))
    info = np.zeros(8, dtype=np.int64)
    _lib.load().ddb_moment_plan_describe(ex.shape[0], ex.shape[1], _lib.ptr(exf), n_t, None, 0, None, 0, None, 0, _lib.ptr(info))
    return bool(info[0]) and ex.shape[1] + n_t > 256


@pytest.mark.parametrize("n,deg,n_t,m", [(12, 3, 12, 9000), (20, 3, 20, 5000), (9, 4, 2, 5001), (40, 2, 5, 7000), (7, 4, 2, 5001),
                                           (30, 2, 3, 7000), (23, 2, 40, 4100), (20, 2, 100, 4100), (12, 2, 150, 4100)])
def test_moment_schedule_other_libraries(ctx, n, deg, n_t, m):
    """Higher degrees (P and Q are degree-3 / degree-4 monomials), n_t != n, a short target block; libraries of a
    few blocks where the staircase saves nothing keep the plain schedule, exactly when the host-only plan says so."""
    rng = np.random.default_rng(n * 10 + deg)
    X = 0.7 * rng.standard_normal((n, m))
    Y = rng.standard_normal((n_t, m))
    ex = oracle.polynomial_exponents(n, deg)
    _res, t1, t0 = _check(ctx, ex, X, Y, tol=2e-13)
    assert t0["gram_schedule"] != 2
    used = _plan_says_used(ex, n_t)
    assert (t1["gram_schedule"] == 2) == used
    if used:
        assert t1["gram_dmma_flops"] < t0["gram_dmma_flops"]
    if (n, deg) in ((20, 3), (40, 2)):
        assert used  # these clearly pay off


def test_moment_schedule_incomplete_library(ctx):
    """A library that is not closed under the factor split: no constant term, every third term removed and a few
    degree-3 terms appended.  The virtual rows / columns are then monomials outside the library; the schedule
    either handles it or steps aside for the plain one -- the blocks must be right in both cases."""
    n, m = 40, 6000
    rng = np.random.default_rng(3)
    ex = oracle.polynomial_exponents(n, 2)
    keep = [j for j in range(1, ex.shape[1]) if j % 3 != 0]
    extra = np.zeros((n, 5), dtype=ex.dtype)
    for c in range(5):
        extra[c, c] = 2
        extra[c + 7, c] = 1
    ex2 = np.concatenate([ex[:, keep], extra], axis=1)
    X = 0.8 * rng.standard_normal((n, m))
    Y = rng.standard_normal((n, m))
    _check(ctx, ex2, X, Y, tol=2e-13)


def test_moment_schedule_mixed_degree_and_no_constant_libraries(ctx):
    """The C3 library plus three cubic terms (cuts by actual degree, entries with a cubic term keep their own two
    terms) and the C3 library without the constant (the constant becomes a virtual term outside the library):
    both run on the moment schedule."""
    n, m = 64, 3000
    rng = np.random.default_rng(8)
    X = 0.8 * rng.standard_normal((n, m))
    Y = rng.standard_normal((n, m))
    ex = oracle.polynomial_exponents(n, 2)
    extra = np.zeros((n, 3), dtype=ex.dtype)
    extra[0, 0] = 3
    extra[1, 1], extra[2, 1] = 2, 1
    extra[5, 2] = extra[6, 2] = extra[7, 2] = 1
    for lib in (np.concatenate([ex, extra], axis=1), ex[:, 1:]):
        _res, t1, _t0 = _check(ctx, lib, X, Y, tol=2e-13)
        assert t1["gram_schedule"] == 2


def test_moment_schedule_is_bit_reproducible_and_sample_views_work(ctx):
    n, m = 64, 300_000
    ctx.set_library(oracle.polynomial_exponents(n, 2))
    ctx.generate(1, n, m, 0, 0, 0.0)
    (Ga, Ra, ya), t = _gram(ctx, True)
    assert t["gram_launches"] == 2
    for _ in range(3):
        (G, R, yy), _t = _gram(ctx, True)
        np.testing.assert_array_equal(G, Ga)
        np.testing.assert_array_equal(R, Ra)
        np.testing.assert_array_equal(yy, ya)
    # a sample view (train / test split, reference data_processing.jl:29-39): blocks of the two parts add up
    first = 100_003
    ctx.set_sample_range(0, first)
    (G1, R1, y1), _t = _gram(ctx, True)
    ctx.set_sample_range(first, m)
    (G2, R2, y2), _t = _gram(ctx, True)
    ctx.set_sample_range(0, -1)
    sc = np.sqrt(np.outer(np.diag(Ga), np.diag(Ga)))
    assert np.max(np.abs(G1 + G2 - Ga) / sc) < 1e-13
    assert np.max(np.abs(R1 + R2 - Ra)) < 1e-13 * np.sqrt(np.max(np.diag(Ga)) * np.max(ya))
    np.testing.assert_allclose(y1 + y2, ya, rtol=1e-13)


def test_library_change_with_same_size_rebuilds_the_moment_map(ctx):
    """Two libraries with identical (n, k) but different terms: the gather map belongs to the library."""
    n, m = 30, 5000
    rng = np.random.default_rng(11)
    X = 0.9 * rng.standard_normal((n, m))
    Y = rng.standard_normal((n, m))
    ex = oracle.polynomial_exponents(n, 2)
    _check(ctx, ex, X, Y)
    ex2 = ex.copy()
    ex2[:, 5], ex2[:, 200] = ex[:, 200].copy(), ex[:, 5].copy()  # swap two terms
    ex2[0, 17] += 1  # and raise one to degree <= 3
    _check(ctx, ex2, X, Y, tol=2e-13)


@pytest.mark.parametrize("n,n_t,m", [(64, 64, 4096), (64, 64, 50_003), (40, 5, 7001), (30, 30, 2500)])
def test_tensor_pipe_builder_is_bit_identical_to_the_fp64_pipe_builder(ctx, n, n_t, m):
    """Degree-2 libraries form their library values on the tensor pipe (gram_mb_kernel: one DMMA with a one-hot
    operand = 64 single products; targets / degree-1 terms are copied).  Every value is the same correctly rounded
    product the FP64-pipe builder (DDB_GRAM_MB=0: gram_kernel's multiplies) forms, and the contraction order is
    unchanged, so the packed blocks must agree bit for bit."""
    rng = np.random.default_rng(n + m)
    X = rng.standard_normal((n, m)) * 3.0
    Y = rng.standard_normal((n_t, m))
    ctx.set_library(oracle.polynomial_exponents(n, 2))
    ctx.upload(X, Y)
    old = os.environ.get("DDB_GRAM_MB")
    try:
        os.environ["DDB_GRAM_MB"] = "1"
        (G1, R1, y1), _ = _gram(ctx, True)
        os.environ["DDB_GRAM_MB"] = "0"
        (G0, R0, y0), _ = _gram(ctx, True)
    finally:
        if old is None:
            del os.environ["DDB_GRAM_MB"]
        else:
            os.environ["DDB_GRAM_MB"] = old
    np.testing.assert_array_equal(G1, G0)
    np.testing.assert_array_equal(R1, R0)
    np.testing.assert_array_equal(y1, y0)
