"""GPU tests of the power-table builder of the single-pair Gram kernel (csrc/gram.cu: gram_pow_kernel): libraries of at
most 64 augmented rows over at most 3 states with degree >= 3 (config C2: Lorenz, degree 5) form every library value as
one power per state, x^a y^b z^c, from a per-stage table of powers -- min(n, degree) factors instead of `degree`.  The
association of mixed terms differs from the factor-by-factor product, so the packed [G | R | yy] agrees with
gram_kernel<64, NF> and with the oracle (Theta Theta', Theta Y', diag Y Y' of reference STLSQ.jl:90-91 on the
materialised library) to rounding, not bit for bit."""
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


def _gram(ctx, power):
    old = os.environ.get("DDB_GRAM_POW")
    os.environ["DDB_GRAM_POW"] = "1" if power else "0"
    try:
        ctx.gram()
        return ctx.gram_download(), ctx.timings()["gram_schedule"]
    finally:
        if old is None:
            del os.environ["DDB_GRAM_POW"]
        else:
            os.environ["DDB_GRAM_POW"] = old


# m = 5 / 16 / 17 / 33 / 65: one to five 16-sample stages (the pipeline runs four stages ahead), ragged last stages
@pytest.mark.parametrize("n,deg,n_t,m", [(3, 5, 3, 100_003), (3, 5, 3, 5), (3, 5, 3, 16), (3, 5, 1, 17), (3, 5, 3, 33),
                                          (3, 5, 3, 65), (3, 3, 3, 5000), (2, 8, 2, 7001), (1, 7, 1, 3000), (3, 4, 6, 31),
                                          (2, 3, 2, 1_000_000)])
def test_power_table_builder_matches_factor_builder_and_oracle(ctx, n, deg, n_t, m):
    rng = np.random.default_rng(n * 100 + deg + m)
    X = rng.standard_normal((n, m)) * 1.3
    Y = rng.standard_normal((n_t, m))
    ex = oracle.polynomial_exponents(n, deg)
    assert ex.shape[1] + n_t <= 64
    ctx.set_library(ex)
    ctx.upload(X, Y)
    (G1, R1, y1), s1 = _gram(ctx, True)
    (G0, R0, y0), s0 = _gram(ctx, False)
    assert s1 == 3 and s0 in (0, 4)  # the power-table kernel really ran
    Th = oracle.evaluate_library(ex, X)
    Gr, Rr = Th @ Th.T, Th @ Y.T
    sc = np.sqrt(np.outer(np.diag(Gr), np.diag(Gr)))
    assert np.max(np.abs(G1 - Gr) / sc) < 1e-13
    assert np.max(np.abs(G1 - G0) / sc) < 1e-13
    rs = np.sqrt(np.max(np.diag(Gr)) * np.max((Y * Y).sum(1)))
    assert np.max(np.abs(R1 - Rr)) < 1e-13 * rs and np.max(np.abs(R1 - R0)) < 1e-13 * rs
    np.testing.assert_allclose(y1, (Y * Y).sum(1), rtol=1e-13)
    np.testing.assert_array_equal(G1, G1.T)
    assert G1[0, 0] == m
    (G2, R2, y2), _ = _gram(ctx, True)  # bit-reproducible from call to call
    np.testing.assert_array_equal(G1, G2)
    np.testing.assert_array_equal(R1, R2)


def test_power_table_builder_steps_aside(ctx):
    """More than 3 states, degree 2, or too many targets for the shared-memory budget: the factor builder runs."""
    rng = np.random.default_rng(3)
    for n, deg, n_t in [(5, 3, 5), (3, 2, 3), (6, 2, 1)]:
        X = rng.standard_normal((n, 4000))
        Y = rng.standard_normal((n_t, 4000))
        ex = oracle.polynomial_exponents(n, deg)
        ctx.set_library(ex)
        ctx.upload(X, Y)
        (G1, R1, _), s1 = _gram(ctx, True)
        assert s1 in (0, 4)
        Th = oracle.evaluate_library(ex, X)
        Gr = Th @ Th.T
        assert np.max(np.abs(G1 - Gr) / np.sqrt(np.outer(np.diag(Gr), np.diag(Gr)))) < 1e-13


def test_sparse_library_without_pure_powers(ctx):
    """A hand-picked library (no constant, mixed terms only, a repeated state): the table covers every power a term
    can ask for, whether or not the pure power is a library term itself."""
    rng = np.random.default_rng(5)
    m = 9000
    X = rng.standard_normal((3, m))
    Y = rng.standard_normal((2, m))
    ex = np.array([[2, 1, 0], [0, 3, 1], [1, 1, 1], [4, 0, 1], [0, 0, 5], [1, 0, 0]], dtype=np.uint8).T.copy()
    ctx.set_library(ex)
    ctx.upload(X, Y)
    (G1, R1, _), s1 = _gram(ctx, True)
    assert s1 == 3
    Th = oracle.evaluate_library(ex, X)
    Gr, Rr = Th @ Th.T, Th @ Y.T
    assert np.max(np.abs(G1 - Gr) / np.sqrt(np.outer(np.diag(Gr), np.diag(Gr)))) < 1e-13
    assert np.max(np.abs(R1 - Rr)) < 1e-13 * np.sqrt(np.max(np.diag(Gr)) * np.max((Y * Y).sum(1)))
