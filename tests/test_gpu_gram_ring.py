"""GPU tests of the ring schedule of the Gram stage (producer SMs -> L2 ring -> TMA + DMMA consumer SMs,
csrc/gram_ring.cu): it must agree with the oracle, with the self-building schedule (csrc/gram.cu) and with
itself bit for bit -- a race between producers and consumers shows up as run-to-run differences."""
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


def _gram(ctx, ring):
    """Plain (non-moment) schedules only: the moment schedule (tests/test_gpu_gram_moment.py) would take
    these libraries first."""
    old = {key: os.environ.get(key) for key in ("DDB_GRAM_RING", "DDB_GRAM_MOMENT")}
    os.environ["DDB_GRAM_RING"] = "1" if ring else "0"
    os.environ["DDB_GRAM_MOMENT"] = "0"
    try:
        ctx.gram()
        return ctx.gram_download(), ctx.timings()["gram_launches"]
    finally:
        for key, val in old.items():
            if val is None:
                del os.environ[key]
            else:
                os.environ[key] = val


@pytest.mark.parametrize("m", [4096, 20_001, 65_536 + 17])
def test_ring_matches_oracle_and_self_building_schedule(ctx, m):
    n = 64
    rng = np.random.default_rng(m)
    X = rng.standard_normal((n, m))
    Y = rng.standard_normal((n, m))
    ex = oracle.polynomial_exponents(n, 2)
    ctx.set_library(ex)
    ctx.upload(X, Y)
    (G1, R1, y1), l1 = _gram(ctx, True)
    (G0, R0, y0), l0 = _gram(ctx, False)
    assert l1 == 3 and l0 == 2  # ring + diagonal pairs + reduce versus self-building + reduce
    Th = oracle.evaluate_library(ex, X)
    Gr, Rr, yr = Th @ Th.T, Th @ Y.T, (Y * Y).sum(1)
    sc = np.sqrt(np.outer(np.diag(Gr), np.diag(Gr)))
    for G, R, yy in ((G1, R1, y1), (G0, R0, y0)):
        assert np.max(np.abs(G - Gr) / sc) < 1e-13
        assert np.max(np.abs(R - Rr)) < 1e-13 * np.sqrt(np.max(np.diag(Gr)) * np.max(yr))
        np.testing.assert_allclose(yy, yr, rtol=1e-13)
        np.testing.assert_array_equal(G, G.T)


def test_ring_is_bit_reproducible_at_scale(ctx):
    """2 x 10^6 samples (31 250 producer windows, ~2000 recyclings of every global ring slot), five runs."""
    n, m = 64, 2_000_000
    ctx.set_library(oracle.polynomial_exponents(n, 2))
    ctx.generate(1, n, m, 0, 0, 0.0)
    (Ga, Ra, ya), launches = _gram(ctx, True)
    assert launches == 3
    for _ in range(4):
        (G, R, yy), _l = _gram(ctx, True)
        np.testing.assert_array_equal(G, Ga)
        np.testing.assert_array_equal(R, Ra)
        np.testing.assert_array_equal(yy, ya)
    (G0, R0, y0), _l = _gram(ctx, False)
    sc = np.sqrt(np.outer(np.diag(G0), np.diag(G0)))
    assert np.max(np.abs(Ga - G0) / sc) < 1e-12
    assert np.max(np.abs(Ra - R0)) < 1e-12 * np.sqrt(np.max(np.diag(G0)) * np.max(y0))
    assert Ga[0, 0] == m  # constant term: a sum of ones is exact


@pytest.mark.parametrize("n,deg,m", [(64, 2, 700_001), (3, 5, 300_000)])
def test_upload_gram_chunked_overlap_matches_plain(ctx, n, deg, m):
    """ddb_upload_gram (chunked upload overlapped by the Gram kernels, per-chunk blocks added in order)
    against ddb_upload + ddb_gram on the same host data."""
    import torch

    rng = np.random.default_rng(5)
    X = np.asfortranarray(rng.standard_normal((n, m)))
    Y = np.asfortranarray(rng.standard_normal((n, m)))
    ctx.set_library(oracle.polynomial_exponents(n, deg))
    ctx.upload(X, Y)
    ctx.gram()
    G0, R0, y0 = ctx.gram_download()
    # pinned host memory -> real overlap
    Xp = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()  # (m, n) row-major == n x m column-major
    Yp = torch.from_numpy(np.ascontiguousarray(Y.T)).pin_memory()
    ctx.upload_gram_raw(Xp.data_ptr(), Yp.data_ptr(), n, m)
    G1, R1, y1 = ctx.gram_download()
    sc = np.sqrt(np.outer(np.diag(G0), np.diag(G0)))
    assert np.max(np.abs(G1 - G0) / sc) < 1e-13
    assert np.max(np.abs(R1 - R0)) < 1e-13 * np.sqrt(np.max(np.diag(G0)) * np.max(y0))
    np.testing.assert_allclose(y1, y0, rtol=1e-13)
    Xd, Yd = ctx.download_data()  # the data stay resident, whole
    np.testing.assert_array_equal(Xd, X)
    np.testing.assert_array_equal(Yd, Y)
    ctx.upload_gram_raw(Xp.data_ptr(), Yp.data_ptr(), n, m)  # deterministic
    G2, _, _ = ctx.gram_download()
    np.testing.assert_array_equal(G1, G2)
    # pageable host memory (what a Julia Matrix or a NumPy array is): the library's own bounce ring, filled by a
    # thread pool -- same chunks, so the blocks are bit-identical, and the resident data are the caller's
    Xh, Yh = np.ascontiguousarray(X.T), np.ascontiguousarray(Y.T)
    ctx.upload_gram_raw(Xh.ctypes.data, Yh.ctypes.data, n, m)
    G3, R3, y3 = ctx.gram_download()
    np.testing.assert_array_equal(G1, G3)
    np.testing.assert_array_equal(R1, R3)
    np.testing.assert_array_equal(y1, y3)
    Xd, Yd = ctx.download_data()
    np.testing.assert_array_equal(Xd, X)
    np.testing.assert_array_equal(Yd, Y)
    # large uploads (2 GB and more; here forced) start with two short chunks so that less of the first copy is exposed
    import os
    os.environ["DDB_UPLOAD_FRONT_MB"] = "1"
    try:
        for px, py in ((Xp.data_ptr(), Yp.data_ptr()), (Xh.ctypes.data, Yh.ctypes.data)):
            ctx.upload_gram_raw(px, py, n, m)
            G4, R4, y4 = ctx.gram_download()
            assert np.max(np.abs(G4 - G0) / sc) < 1e-13
            assert np.max(np.abs(R4 - R0)) < 1e-13 * np.sqrt(np.max(np.diag(G0)) * np.max(y0))
            np.testing.assert_allclose(y4, y0, rtol=1e-13)
            Xd, Yd = ctx.download_data()
            np.testing.assert_array_equal(Xd, X)
            np.testing.assert_array_equal(Yd, Y)
    finally:
        del os.environ["DDB_UPLOAD_FRONT_MB"]


@pytest.mark.parametrize("n,n_t", [(65, 65), (66, 66), (70, 3), (64, 1), (72, 72)])
def test_ring_plan_variants(ctx, n, n_t):
    """Library sizes around the ring schedule's planning cases: with and without a short last block row (strip
    consumers), last block rows of 40 / 66 / 100 live rows, more block pairs than consumers."""
    m = 4500
    rng = np.random.default_rng(n * 100 + n_t)
    X = rng.standard_normal((n, m))
    Y = rng.standard_normal((n_t, m))
    ex = oracle.polynomial_exponents(n, 2)
    ctx.set_library(ex)
    ctx.upload(X, Y)
    (G1, R1, y1), l1 = _gram(ctx, True)
    assert l1 == 3  # the ring schedule ran
    Th = oracle.evaluate_library(ex, X)
    Gr, Rr, yr = Th @ Th.T, Th @ Y.T, (Y * Y).sum(1)
    sc = np.sqrt(np.outer(np.diag(Gr), np.diag(Gr)))
    assert np.max(np.abs(G1 - Gr) / sc) < 1e-13
    assert np.max(np.abs(R1 - Rr)) < 1e-13 * np.sqrt(np.max(np.diag(Gr)) * np.max(yr))
    np.testing.assert_allclose(y1, yr, rtol=1e-13)
    np.testing.assert_array_equal(G1, G1.T)
