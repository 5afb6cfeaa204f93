"""GPU tests of the four-warp kernel for single-pair libraries (csrc/gram.cu: gram_w4_kernel, k + n_t <= 64: config C2
and the README problem): the 36 live sub-tiles of the 64 x 64 diagonal pair are dealt 9 + 9 + 9 + 9 to four warps by
sub-tile rows.  Same products, same accumulation order per entry, same work plan as gram_kernel<64, NF>: the packed
[G | R | yy] must agree with it bit for bit, and with the oracle (Theta Theta', Theta Y', diag Y Y' of reference
STLSQ.jl:90-91 on the materialised library) to rounding."""
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


def _gram(ctx, w4):
    old = os.environ.get("DDB_GRAM_W4")
    oldp = os.environ.get("DDB_GRAM_POW")
    os.environ["DDB_GRAM_W4"] = "1" if w4 else "0"
    os.environ["DDB_GRAM_POW"] = "0"  # the factor-by-factor builder on both layouts
    try:
        ctx.gram()
        return ctx.gram_download(), ctx.timings()["gram_schedule"]
    finally:
        if oldp is None:
            del os.environ["DDB_GRAM_POW"]
        else:
            os.environ["DDB_GRAM_POW"] = oldp
        if old is None:
            del os.environ["DDB_GRAM_W4"]
        else:
            os.environ["DDB_GRAM_W4"] = old


@pytest.mark.parametrize("n,deg,n_t,m", [(3, 5, 3, 100_003), (3, 5, 3, 5), (3, 5, 3, 16), (3, 5, 1, 17), (3, 5, 3, 33),
                                          (3, 5, 8, 20_000), (3, 3, 3, 5000), (2, 8, 2, 7001), (1, 7, 1, 3000),
                                          (5, 2, 5, 4096), (6, 2, 1, 999), (2, 1, 7, 64), (9, 2, 9, 1_000_001)])
def test_four_warp_kernel_is_bit_identical_to_the_two_warp_kernel(ctx, n, deg, n_t, m):
    rng = np.random.default_rng(n * 100 + deg + m)
    X = rng.standard_normal((n, m)) * 1.3
    Y = rng.standard_normal((n_t, m))
    ex = oracle.polynomial_exponents(n, deg)
    assert ex.shape[1] + n_t <= 64
    ctx.set_library(ex)
    ctx.upload(X, Y)
    (G1, R1, y1), s1 = _gram(ctx, True)
    (G0, R0, y0), s0 = _gram(ctx, False)
    assert s1 == 4 and s0 == 0
    np.testing.assert_array_equal(G1, G0)
    np.testing.assert_array_equal(R1, R0)
    np.testing.assert_array_equal(y1, y0)
    Th = oracle.evaluate_library(ex, X)
    Gr, Rr = Th @ Th.T, Th @ Y.T
    sc = np.sqrt(np.outer(np.diag(Gr), np.diag(Gr)))
    assert np.max(np.abs(G1 - Gr) / sc) < 1e-13
    assert np.max(np.abs(R1 - Rr)) < 1e-13 * np.sqrt(np.max(np.diag(Gr)) * np.max((Y * Y).sum(1)))
    np.testing.assert_allclose(y1, (Y * Y).sum(1), rtol=1e-13)
    np.testing.assert_array_equal(G1, G1.T)
    assert G1[0, 0] == m
