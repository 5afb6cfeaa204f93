"""Host logic of the moment schedule of the Gram stage (csrc/gram_moment.cu) through the host-only entry point
ddb_moment_plan_describe: no GPU, no compute calls.  The GPU parity of the schedule is tests/test_gpu_gram_moment.py."""
import numpy as np
import pytest

import oracle
from ddb200 import _lib


def _moment_plan(ex, n_t):
    lib = _lib.load()
    n, k = ex.shape
    exf = np.asfortranarray(ex.astype(np.uint8))
    info = np.zeros(8, dtype=np.int64)
    count = lib.ddb_moment_plan_describe(n, k, _lib.ptr(exf), n_t, None, 0, None, 0, None, 0, _lib.ptr(info))
    used, nrows, ncols, nrb, ncb, npairs, nvirt, deg = info.tolist()
    if not used:
        assert count == 0
        return None
    assert count == k * k + k * n_t + n_t
    src = np.zeros(count, dtype=np.int32)
    tab = np.zeros(((nrb + ncb) * 128, 8), dtype=np.uint16)
    pairs = np.zeros((npairs, 3), dtype=np.int32)
    assert lib.ddb_moment_plan_describe(n, k, _lib.ptr(exf), n_t, _lib.ptr(src), src.size, _lib.ptr(tab), tab.size,
                                        _lib.ptr(pairs), pairs.shape[0], None) == count
    return dict(src=src, tab=tab, pairs=pairs, nrb=nrb, ncb=ncb, npairs=npairs, nrows=nrows, ncols=ncols, nvirt=nvirt)


def _virtual_exponents(tab, n, n_t):
    """Exponent vector over [states | targets] of every virtual term of the factor table."""
    out = np.zeros((tab.shape[0], n + n_t), dtype=np.int64)
    for r in range(tab.shape[0]):
        for f in tab[r]:
            if f == 0xFFFF:
                continue
            out[r, (n + (f & 0x7FFF)) if (f & 0x8000) else f] += 1
    return out


@pytest.mark.parametrize("n,deg,n_t", [(64, 2, 64), (20, 3, 20), (9, 4, 2), (40, 2, 5), (12, 3, 12)])
def test_moment_plan_maps_every_packed_entry_to_its_moment(n, deg, n_t):
    """Complete polynomial libraries: every entry is gathered from its moment, one source per distinct moment."""
    _check_plan(oracle.polynomial_exponents(n, deg), n_t, complete=True)


def _check_plan(ex, n_t, complete):
    """For EVERY entry of the packed [G | R | yy] the product of the virtual row term and the virtual column term
    it is gathered from is the product monomial of the entry.  Checked on exponent vectors: all entries through
    a random linear form of the exponents (a 60-bit fingerprint), 200 000 random entries exactly.  For complete
    libraries entries of the same moment share one source and different moments have different sources; the
    block pairs are exactly the ones the entries touch."""
    n = ex.shape[0]
    k = ex.shape[1]
    plan = _moment_plan(ex, n_t)
    assert plan is not None
    src, pairs, nrb = plan["src"], plan["pairs"], plan["nrb"]
    vex = _virtual_exponents(plan["tab"], n, n_t)
    pair, off = src >> 14, src & 16383
    row = pairs[pair, 0] * 128 + off // 128
    col = pairs[pair, 1] * 128 + off % 128
    assert ((pairs[:, 0] < nrb) != (pairs[:, 1] < nrb)).all()  # a row block and a column block, either order
    assert (pairs[pairs[:, 2] == 128, 0] < nrb).all()
    # pairs at the edge of the staircase that only use their first 64 / 32 rows come last and are marked
    assert set(pairs[:, 2].tolist()) <= {32, 64, 128} and (np.diff(pairs[:, 2]) <= 0).all()
    assert ((off // 128) < pairs[pair, 2]).all()
    top = np.zeros(plan["npairs"], dtype=np.int64)  # largest row / column used per pair: the class is the tightest
    np.maximum.at(top, pair, off // 128)
    left = np.zeros(plan["npairs"], dtype=np.int64)
    np.maximum.at(left, pair, off % 128)
    tight = np.minimum(top, left)  # a pair may be run either way round
    np.testing.assert_array_equal(np.where(tight < 32, 32, np.where(tight < 64, 64, 128)), pairs[:, 2])
    full = np.concatenate([ex.T.astype(np.int64), np.zeros((k, n_t), dtype=np.int64)], axis=1)  # library terms
    tgt = np.concatenate([np.zeros((n_t, n), dtype=np.int64), np.eye(n_t, dtype=np.int64)], axis=1)  # targets
    rng = np.random.default_rng(0)
    w = rng.integers(1, 1 << 56, size=n + n_t, dtype=np.int64)
    hv, hf, ht = vex @ w, full @ w, tgt @ w
    got = hv[row] + hv[col]
    want = np.concatenate([(hf[None, :] + hf[:, None]).ravel(), (ht[:, None] + hf[None, :]).ravel(), 2 * ht])
    np.testing.assert_array_equal(got, want)
    # exact check of a sample of entries (entry e of G is (t, u) = (e % k, e // k); of R (t, j))
    e = rng.integers(0, k * k + k * n_t, size=200_000)
    a = np.where(e < k * k, e % k, (e - k * k) % k)
    b = np.where(e < k * k, e // k, (e - k * k) // k)
    second = np.where((e < k * k)[:, None], full[np.minimum(b, k - 1)], tgt[np.minimum(b, n_t - 1)])
    np.testing.assert_array_equal(vex[row[e]] + vex[col[e]], full[a] + second)
    # one source per distinct moment
    G_src = src[: k * k].reshape(k, k)
    np.testing.assert_array_equal(G_src, G_src.T)
    if complete:
        keys, inv = np.unique(want[: k * k], return_inverse=True)
        rep = np.zeros(keys.size, dtype=np.int64)
        rep[inv] = src[: k * k]
        np.testing.assert_array_equal(rep[inv], src[: k * k])
        assert np.unique(rep).size == keys.size
    assert np.unique(pair).size == plan["npairs"]  # no block pair without a needed entry
    nblk = -(-(k + n_t) // 128)
    cost = (pairs[:, 2] == 128).sum() + 0.53 * (pairs[:, 2] == 64).sum() + 0.45 * (pairs[:, 2] == 32).sum()
    assert cost <= 0.9 * (0.5 * nblk * (nblk - 1) + 0.6 * nblk)
    return plan


def test_moment_plan_lorenz96_counts():
    """Config C3: 2145 virtual rows and columns (the library is closed under the split), 78 block pairs instead
    of 153 off-diagonal + 18 diagonal ones."""
    plan = _moment_plan(oracle.polynomial_exponents(64, 2), 64)
    assert (plan["nrows"], plan["ncols"], plan["nrb"], plan["ncb"], plan["npairs"]) == (2145, 2145, 18, 18, 78)
    # 62 whole pairs, 6 of 64 rows (cost 0.53), 10 of 32 rows (cost 0.45): 69.7 pair units
    assert [(plan["pairs"][:, 2] == r).sum() for r in (128, 64, 32)] == [62, 6, 10]


def test_moment_plan_of_mixed_degree_and_incomplete_libraries():
    """Libraries that are not closed under the cut in the middle of the padded factor list.  (1) The C3 library
    plus three cubic terms: the padded cut of two quadratic terms would be a linear and a cubic monomial, so the
    lower / upper halves of the actual factors are used, and entries with a cubic term keep their own two terms;
    no monomial outside the library appears.  (2) The C3 library without its constant term: the constant becomes
    the one virtual term outside the library."""
    ex = oracle.polynomial_exponents(64, 2)
    extra = np.zeros((64, 3), dtype=ex.dtype)
    extra[0, 0] = 3
    extra[1, 1], extra[2, 1] = 2, 1
    extra[5, 2] = extra[6, 2] = extra[7, 2] = 1
    mixed = np.concatenate([ex, extra], axis=1)
    plan = _check_plan(mixed, 64, complete=False)
    assert (plan["nrows"], plan["ncols"]) == (mixed.shape[1], mixed.shape[1]) and plan["npairs"] <= 82
    plan = _check_plan(ex[:, 1:], 64, complete=False)
    assert plan["nrows"] == ex.shape[1] and plan["npairs"] <= 78


def test_moment_plan_steps_aside():
    """Libraries where the staircase saves nothing keep the plain schedule."""
    assert _moment_plan(oracle.polynomial_exponents(12, 2), 150) is None  # many targets: 4 block pairs against 1 + 2 diagonal


def _wave_plan(ex, n_t, m, nctas=148):
    lib = _lib.load()
    n, k = ex.shape
    exf = np.asfortranarray(ex.astype(np.uint8))
    info = np.zeros(4, dtype=np.float64)
    cnt = lib.ddb_moment_wave_describe(n, k, _lib.ptr(exf), n_t, m, nctas, None, 0, _lib.ptr(info))
    assert cnt > 0
    seg = np.zeros((cnt, 7), dtype=np.int64)
    assert lib.ddb_moment_wave_describe(n, k, _lib.ptr(exf), n_t, m, nctas, _lib.ptr(seg), cnt, None) == cnt
    return seg, info


@pytest.mark.parametrize("n,deg,n_t,m,nctas", [(64, 2, 64, 10_000_000, 148), (64, 2, 64, 1_250_000, 148), (64, 2, 64, 4097, 148),
                                                (64, 2, 64, 33, 148), (20, 3, 20, 200_000, 148), (100, 2, 100, 1_000_000, 148),
                                                (40, 2, 5, 70_001, 132), (12, 3, 12, 9000, 148)])
def test_moment_wave_plan_covers_every_pair_and_stage_once(n, deg, n_t, m, nctas):
    """Host logic of the lock-step waves: every block pair of the moment plan is covered over all stages exactly
    once by contiguous ranges, one partial-tile slot per segment with the slots of a pair contiguous and in stage
    order (fixed summation order), at most nctas CTAs per wave, and the planned time is within 2 % of a perfect
    split when there is enough work per CTA."""
    ex = oracle.polynomial_exponents(n, deg)
    plan = _moment_plan(ex, n_t)
    seg, info = _wave_plan(ex, n_t, m, nctas)
    nstages = -(-m // 16)
    rows_of = {(int(a), int(b)): int(r) for a, b, r in plan["pairs"]}
    cover = {}
    for cta, bi, bj, slot, b, e, rows in seg.tolist():
        assert 0 <= cta < nctas and 0 <= b <= e <= nstages
        assert rows_of[(bi, bj)] == rows
        cover.setdefault((bi, bj), []).append((b, e, slot))
    assert set(cover) == set(rows_of)
    for ranges in cover.values():
        ranges.sort()
        assert ranges[0][0] == 0 and ranges[-1][1] == nstages
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(len(ranges) - 1))
        assert [r[2] for r in ranges] == list(range(ranges[0][2], ranges[0][2] + len(ranges)))
    assert sorted(seg[:, 3].tolist()) == list(range(seg.shape[0])) and info[3] == seg.shape[0]
    # a CTA takes part in every homogeneous wave at most once; in the final wave most CTAs take one aligned piece of
    # one pair, the rest share the pairs' remainders as a stream cut where the common finishing time falls, so a
    # CTA may change pair a few times there
    per_cta = np.bincount(seg[:, 0], minlength=nctas)
    assert per_cta.max() <= info[2] + 8
    # planned finishing time = the busiest CTA's share, recomputed here from the segments
    cost_of = {128: 1.0, 64: 0.53, 32: 0.45}
    busy = np.zeros(nctas)
    for cta, bi, bj, slot, b, e, rows in seg.tolist():
        busy[cta] += cost_of[rows] * (e - b) / nstages
    assert abs(busy.max() - info[0]) < 1e-9
    if nstages >= 64 * nctas:
        assert info[1] <= info[0] <= 1.02 * info[1]


def test_moment_wave_plan_lorenz96_is_within_one_percent_of_a_perfect_split():
    """C3 at full size: three homogeneous waves of whole pairs (37 x 4 ranges, 21 x 7, 2 x 74), then the last 2 whole,
    6 half and 10 quarter pairs as one stream that fills every CTA to a common finishing time: 69.7 / 148 = 0.471
    pair streams per CTA to a tenth of a percent."""
    seg, info = _wave_plan(oracle.polynomial_exponents(64, 2), 64, 10_000_000)
    assert info[2] <= 4 and info[0] <= 1.001 * info[1]


def test_moment_plan_random_libraries():
    """Random sparse libraries (duplicate terms, missing constants, mixed degrees, a single term, all-constant
    columns): whenever the analysis accepts a library its gather map must be right for every packed entry."""
    rng = np.random.default_rng(0)
    libs = []
    for _ in range(25):
        n, k, n_t, maxdeg = int(rng.integers(1, 12)), int(rng.integers(1, 700)), int(rng.integers(1, 40)), int(rng.integers(1, 6))
        ex = np.zeros((n, k), dtype=np.uint8)
        for j in range(k):
            for _f in range(int(rng.integers(0, maxdeg + 1))):
                ex[rng.integers(0, n), j] += 1
        libs.append((ex, n_t))
    libs += [(np.zeros((3, 300), dtype=np.uint8), 5), (np.ones((2, 400), dtype=np.uint8), 2), (np.zeros((1, 1), dtype=np.uint8), 1)]
    used = 0
    for ex, n_t in libs:
        n, k = ex.shape
        plan = _moment_plan(ex, n_t)
        if plan is None:
            continue
        used += 1
        src, pairs = plan["src"], plan["pairs"]
        vex = _virtual_exponents(plan["tab"], n, n_t)
        pair, off = src >> 14, src & 16383
        row = pairs[pair, 0] * 128 + off // 128
        col = pairs[pair, 1] * 128 + off % 128
        full = np.concatenate([ex.T.astype(np.int64), np.zeros((k, n_t), dtype=np.int64)], axis=1)
        tgt = np.concatenate([np.zeros((n_t, n), dtype=np.int64), np.eye(n_t, dtype=np.int64)], axis=1)
        w = rng.integers(1, 1 << 56, size=n + n_t, dtype=np.int64)
        hv, hf, ht = vex @ w, full @ w, tgt @ w
        want = np.concatenate([(hf[None, :] + hf[:, None]).ravel(), (ht[:, None] + hf[None, :]).ravel(), 2 * ht])
        np.testing.assert_array_equal(hv[row] + hv[col], want)
        assert ((off // 128) < pairs[pair, 2]).all()
    assert used >= 15


# ---------------------------------------------------------------------------------------------------------------
# Tensor-pipe builder tables (ddb_moment_builder_describe): the ops of every block pair, replayed on the host
# ---------------------------------------------------------------------------------------------------------------
_OPS, _COPIES = 11, 5                      # kMbOpSlots, kMbCopySlots (csrc/mma_common.cuh)
_WARP_WORDS = _OPS * 64 + _COPIES * 32


def _builder_tables(ex, n_t, npairs, nblocks):
    lib = _lib.load()
    n, k = ex.shape
    exf = np.asfortranarray(ex.astype(np.uint8))
    optab = np.zeros((npairs, 8, _WARP_WORDS), dtype=np.uint32)
    opcnt = np.zeros((npairs, 8), dtype=np.uint8)
    block_ops = np.zeros(nblocks, dtype=np.int32)
    got = lib.ddb_moment_builder_describe(n, k, _lib.ptr(exf), n_t, _lib.ptr(optab), optab.size, _lib.ptr(opcnt), opcnt.size,
                                          _lib.ptr(block_ops), block_ops.size)
    return got, optab, opcnt, block_ops


def _replay_builder(optab, opcnt, raw):
    """What gram_mb_kernel's ops leave in a Theta stage (2 tiles x 16 samples x 132 columns): every product op is one
    DMMA m8n8k4 with A[g][t] and B[t][g] read through the per-lane offsets, lane (g, t) stores D[g][2t], D[g][2t+1];
    every copy slot moves one double per lane."""
    stage = np.zeros(2 * 16 * 132)
    lanes = np.arange(32)
    g, t = lanes >> 2, lanes & 3
    nops = 0
    for w in range(8):
        ops = optab[w, :_OPS * 64].reshape(_OPS, 32, 2)
        cps = optab[w, _OPS * 64:].reshape(_COPIES, 32)
        for q in range(_OPS):
            kk, slot = q // 3, q % 3
            if slot == 2 and not (opcnt[w] >> kk) & 1:
                continue
            e = ops[q]
            offa, offb = e[:, 0] & 0xFFFF, e[:, 0] >> 16
            d0, d1 = e[:, 1] & 0xFFFF, e[:, 1] >> 16
            assert not (offa % 8).any() and not (offb % 8).any() and not (d0 % 8).any() and not (d1 % 8).any()
            A = np.zeros((8, 4))
            B = np.zeros((4, 8))
            A[g, t] = raw[offa // 8]
            B[t, g] = raw[offb // 8]
            assert ((A != 0).astype(int) @ (B != 0).astype(int) <= 1).all()  # one-hot operand: every output is a single product
            D = A @ B
            stage[d0 // 8] = D[g, 2 * t]
            stage[d1 // 8] = D[g, 2 * t + 1]
            nops += 1
        for q in range(int(opcnt[w]) >> 4):
            c = cps[q]
            stage[(c >> 16) // 8] = raw[(c & 0xFFFF) // 8]
    return stage.reshape(2, 16, 132), nops


@pytest.mark.parametrize("n,deg,n_t", [(64, 2, 64), (40, 2, 5), (30, 2, 30)])
def test_builder_tables_form_every_virtual_term(n, deg, n_t):
    """Degree-2 libraries: replaying the op tables of every block pair on random samples fills both Theta tiles with
    the virtual terms' values (bitwise: each is a single product), unused rows of half / quarter pairs excepted."""
    ex = oracle.polynomial_exponents(n, deg)
    plan = _moment_plan(ex, n_t)
    assert plan is not None
    nblocks = plan["nrb"] + plan["ncb"]
    got, optab, opcnt, block_ops = _builder_tables(ex, n_t, plan["npairs"], nblocks)
    assert got == plan["npairs"]
    assert block_ops.min() >= 0 and block_ops.max() <= 16
    rng = np.random.default_rng(5)
    X = rng.standard_normal((16, n))
    Y = rng.standard_normal((16, n_t))
    px = n + (4 - n % 16) % 16  # mb_pitch: the X tile's rows are padded to a pitch = 4 (mod 16)
    Xp = np.full((16, px), np.nan)
    Xp[:, :n] = X
    raw = np.concatenate([np.ones(16), [0.0, 0.0], Xp.ravel(), Y.ravel()])
    # value of every virtual term on the 16 samples (product of its factor codes; missing factors = 1)
    tab = plan["tab"]
    vals = np.ones((tab.shape[0], 16))
    nfac = np.zeros(tab.shape[0], dtype=int)
    for r in range(tab.shape[0]):
        for f in tab[r]:
            if f == 0xFFFF:
                continue
            vals[r] = vals[r] * (Y[:, f & 0x7FFF] if (f & 0x8000) else X[:, f])
            nfac[r] += 1
    row_terms, col_terms = plan["nrows"] + n_t, n_t + plan["ncols"]
    for p in range(plan["npairs"]):
        b0, b1, used = plan["pairs"][p]
        stage, nops = _replay_builder(optab[p], opcnt[p], raw)
        assert 64 <= nops <= 8 * _OPS
        for tile, blk in enumerate((b0, b1)):
            for c in range(128):
                idx = blk * 128 + c
                valid = idx < row_terms if blk < plan["nrb"] else idx - plan["nrb"] * 128 < col_terms
                if tile == 0 and c >= used:
                    continue
                want = vals[idx] if valid else np.zeros(16)
                assert np.array_equal(stage[tile, :, c], want), (p, tile, c)


def test_builder_tables_absent_for_cubic_libraries():
    ex = oracle.polynomial_exponents(20, 3)
    plan = _moment_plan(ex, 20)
    assert plan is not None
    got, _, _, _ = _builder_tables(ex, 20, plan["npairs"], plan["nrb"] + plan["ncb"])
    assert got == 0


# ------------------------------------------------------------------------------------------------
# power-table builder of the single-pair kernel (csrc/gram.cu: gram_pow_kernel), host logic only
# ------------------------------------------------------------------------------------------------
def _pow_factors(ex, n_t):
    lib = _lib.load()
    n, k = ex.shape
    exf = np.asfortranarray(ex.astype(np.uint8))
    tab = np.full((64, 8), 0xFFFF, dtype=np.uint16)
    deg = np.zeros(1, dtype=np.int32)
    nf = lib.ddb_pow_factors_describe(n, k, _lib.ptr(exf), n_t, _lib.ptr(tab), _lib.ptr(deg))
    return nf, int(deg[0]), tab


@pytest.mark.parametrize("n,deg,n_t", [(3, 5, 3), (2, 8, 2), (1, 7, 1), (3, 3, 6), (3, 4, 1)])
def test_pow_factors_replay(n, deg, n_t):
    """Replays the table on the CPU: every library value is the product of one table-of-powers entry per state with a
    non-zero exponent (powers formed left to right, as the kernel's (sample, state) threads form them), targets are
    single raw values, rows past the library have no factor."""
    ex = oracle.polynomial_exponents(n, deg)
    k = ex.shape[1]
    nf, d, tab = _pow_factors(ex, n_t)
    assert d == deg and nf == min(n, deg)
    rng = np.random.default_rng(n * 10 + deg)
    x = rng.standard_normal(n) * 1.5
    powers = np.empty(n * deg)
    for i in range(n):
        p = x[i]
        powers[i * deg] = p
        for e in range(1, deg):
            p = p * x[i]
            powers[i * deg + e] = p
    for j in range(k):
        codes = [int(c) for c in tab[j] if c != 0xFFFF]
        assert len(codes) == int(np.count_nonzero(ex[:, j])) <= nf
        assert sorted(c // deg for c in codes) == [i for i in range(n) if ex[i, j] > 0]  # one code per state, state order
        for c in codes:
            assert c % deg + 1 == ex[c // deg, j]
        val = 1.0
        for c in codes:
            val *= powers[c]
        ref = float(np.prod(x ** ex[:, j].astype(float)))
        assert abs(val - ref) <= 1e-13 * max(1.0, abs(ref))
    for t in range(n_t):
        assert tab[k + t, 0] == (0x8000 | t) and (tab[k + t, 1:] == 0xFFFF).all()
    assert (tab[k + n_t:] == 0xFFFF).all()


def test_pow_factors_step_aside():
    """More than 3 states, degree below 3, or more than 64 augmented rows: the builder does not apply."""
    for n, deg, n_t in [(5, 3, 5), (3, 2, 3), (6, 2, 1), (3, 5, 9), (4, 3, 2)]:
        nf, d, tab = _pow_factors(oracle.polynomial_exponents(n, deg), n_t)
        assert nf == 0 and d == deg and (tab == 0xFFFF).all()
    # a hand-picked library without pure powers
    ex = np.array([[2, 1, 0], [0, 3, 1], [1, 1, 1], [4, 0, 1], [0, 0, 5], [1, 0, 0]], dtype=np.uint8).T.copy()
    nf, d, tab = _pow_factors(ex, 2)
    assert nf == 3 and d == 5
    assert [int(c) for c in tab[0] if c != 0xFFFF] == [0 * 5 + 1, 1 * 5 + 0]
    assert [int(c) for c in tab[4] if c != 0xFFFF] == [2 * 5 + 4]
