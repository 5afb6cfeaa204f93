"""GPU tests of the prefix-product builder of the single-pair Gram kernel (csrc/gram.cu: gram_tree_kernel): libraries
of at most 64 augmented rows that are closed under "drop the last factor" form every library value as parent x last
factor -- one multiply per value, the same association as the factor-by-factor product of gram_kernel<64, NF>, so
the packed [G | R | yy] must agree with it bit for bit, and with the oracle (Theta Theta', Theta Y', diag Y Y' of
reference STLSQ.jl:90-91 on the materialised library) to rounding."""
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


def _gram(ctx, tree):
    old = os.environ.get("DDB_GRAM_TREE")
    oldp = os.environ.get("DDB_GRAM_POW")
    os.environ["DDB_GRAM_TREE"] = "1" if tree else "0"
    os.environ["DDB_GRAM_POW"] = "0"  # the comparison is with the factor-by-factor builder, not the power-table one
    try:
        ctx.gram()
        return ctx.gram_download()
    finally:
        if oldp is None:
            del os.environ["DDB_GRAM_POW"]
        else:
            os.environ["DDB_GRAM_POW"] = oldp
        if old is None:
            del os.environ["DDB_GRAM_TREE"]
        else:
            os.environ["DDB_GRAM_TREE"] = old


@pytest.mark.parametrize("n,deg,n_t,m", [(3, 5, 3, 100_003), (3, 5, 3, 16), (3, 3, 3, 5000), (2, 8, 2, 7001), (5, 2, 5, 4096),
                                          (6, 2, 1, 999), (3, 5, 8, 20_000), (1, 7, 1, 3000)])
def test_tree_builder_is_bit_identical_to_the_factor_builder(ctx, n, deg, n_t, m):
    rng = np.random.default_rng(n * 100 + deg)
    X = rng.standard_normal((n, m)) * 1.7
    Y = rng.standard_normal((n_t, m))
    ex = oracle.polynomial_exponents(n, deg)
    assert ex.shape[1] + n_t <= 64
    ctx.set_library(ex)
    ctx.upload(X, Y)
    G1, R1, y1 = _gram(ctx, True)
    G0, R0, y0 = _gram(ctx, False)
    np.testing.assert_array_equal(G1, G0)
    np.testing.assert_array_equal(R1, R0)
    np.testing.assert_array_equal(y1, y0)
    Th = oracle.evaluate_library(ex, X)
    Gr, Rr = Th @ Th.T, Th @ Y.T
    sc = np.sqrt(np.outer(np.diag(Gr), np.diag(Gr)))
    assert np.max(np.abs(G1 - Gr) / sc) < 1e-13
    assert np.max(np.abs(R1 - Rr)) < 1e-13 * np.sqrt(np.max(np.diag(Gr)) * np.max((Y * Y).sum(1)))
    assert G1[0, 0] == m


def test_libraries_that_are_not_prefix_closed_keep_the_factor_builder(ctx):
    """No constant / a missing intermediate term / a permuted term order: the programs do not apply (or apply to the
    permuted columns) and the result still matches the oracle."""
    n, m = 3, 6000
    rng = np.random.default_rng(9)
    X = rng.standard_normal((n, m))
    Y = rng.standard_normal((n, m))
    full = oracle.polynomial_exponents(n, 4)
    for ex in (full[:, 1:], np.delete(full, 5, axis=1), full[:, rng.permutation(full.shape[1])]):
        ex = np.ascontiguousarray(ex)
        ctx.set_library(ex)
        ctx.upload(X, Y)
        G1, R1, _ = _gram(ctx, True)
        G0, R0, _ = _gram(ctx, False)
        np.testing.assert_array_equal(G1, G0)
        np.testing.assert_array_equal(R1, R0)
        Th = oracle.evaluate_library(ex, X)
        Gr = Th @ Th.T
        sc = np.sqrt(np.outer(np.diag(Gr), np.diag(Gr)))
        assert np.max(np.abs(G1 - Gr) / sc) < 1e-13
