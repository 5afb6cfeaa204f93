"""CPU tests of the host-side mirror of the reference interface (no GPU, no compute calls)."""
import math

import numpy as np
import pytest

import ddb200
import oracle


@pytest.mark.parametrize("n,d", [(1, 4), (3, 2), (3, 5), (5, 3), (64, 2)])
def test_polynomial_basis_matches_reference_order(n, d):
    np.testing.assert_array_equal(ddb200.polynomial_basis(n, d).exponents, oracle.polynomial_exponents(n, d))


@pytest.mark.parametrize("n,d", [(1, 1), (3, 1), (2, 3)])
def test_monomial_basis_matches_reference_order(n, d):
    np.testing.assert_array_equal(ddb200.monomial_basis(n, d).exponents, oracle.monomial_exponents(n, d))


def test_basis_removes_duplicates_keeps_order():
    # __preprocess_basis / _unique_symbolics! (src/basis/type.jl:91-135, 523-532)
    e = np.array([[0, 1, 1, 2, 0], [0, 0, 0, 1, 0]])
    b = ddb200.Basis(e)
    assert b.exponents.T.tolist() == [[0, 0], [1, 0], [2, 1]]
    assert b.term_strings() == ["1", "x1", "x1^2*x2"]


def test_basis_rejects_non_monomials():
    with pytest.raises(ValueError):
        ddb200.Basis(np.array([[0.5, 1.0]]))
    with pytest.raises(ValueError):
        ddb200.Basis(np.array([[-1, 1]]))


def test_sr3_stores_half_squared_thresholds():
    # SR3.jl:63
    a = ddb200.SR3([0.2, 0.4])
    np.testing.assert_allclose(a.thresholds, [0.02, 0.08])
    b = ddb200.SR3([0.2, 0.4], 1.0, ddb200.SoftThreshold())
    np.testing.assert_allclose(b.thresholds, [0.2, 0.4])


def test_algorithm_constructor_asserts():
    with pytest.raises(AssertionError):
        ddb200.STLSQ([0.1, -0.1])
    with pytest.raises(AssertionError):
        ddb200.STLSQ(0.1, -1.0)
    with pytest.raises(AssertionError):
        ddb200.ADMM(0.1, 0.0)
    with pytest.raises(AssertionError):
        ddb200.SR3(0.1, 0.0)


def test_default_options_match_reference():
    o = ddb200.DataDrivenCommonOptions()
    assert o.maxiters == 100 and o.digits == 10 and o.selector == "bic"
    assert o.abstol == o.reltol == math.sqrt(np.finfo(float).eps)


@pytest.mark.parametrize("kw", [dict(denoise=True), dict(normalize="MinMaxSomething"), dict(roundingmode="RoundNearest"),
                                dict(selector="mdl")])
def test_unsupported_options_raise(kw):
    from ddb200.api import _check_options

    with pytest.raises(ValueError):
        _check_options(ddb200.DataDrivenCommonOptions(**kw))


def test_problem_fit_inputs():
    X = np.arange(12.0).reshape(3, 4)
    p = ddb200.DiscreteDataDrivenProblem(X)
    a, b = p.fit_inputs()
    np.testing.assert_array_equal(a, X[:, :-1])  # get_oop_args drops the final sample (type.jl:448-455)
    np.testing.assert_array_equal(b, X[:, 1:])  # get_implicit_data (type.jl:434)
    c = ddb200.ContinuousDataDrivenProblem(X, DX=2 * X)
    np.testing.assert_array_equal(c.fit_inputs()[1], 2 * X)
    with pytest.raises(ValueError):
        ddb200.ContinuousDataDrivenProblem(X)
    with pytest.raises(ValueError):
        ddb200.DirectDataDrivenProblem(X, np.zeros((1, 5)))
    with pytest.raises(ValueError):
        ddb200.DirectDataDrivenProblem(X, np.full((1, 4), np.nan))


def test_construct_coefficients_matches_oracle():
    rng = np.random.default_rng(0)
    C = rng.standard_normal((3, 7)) * 10.0 ** rng.integers(-12, 3, (3, 7))
    np.testing.assert_array_equal(ddb200.construct_coefficients(C), oracle.construct_coefficients(C))


def test_b200sparse_wraps_only_known_algorithms():
    with pytest.raises(ValueError):
        ddb200.B200Sparse(object())


@pytest.mark.parametrize("m,world", [(1001, 2), (10_000_000, 8), (17, 4), (16, 3)])
def test_shard_ranges_tile_the_samples(m, world):
    lo_prev = 0
    for r in range(world):
        lo, hi = ddb200.dist.shard_range(m, r, world)
        assert lo == lo_prev and hi >= lo
        if r < world - 1:
            assert hi % 16 == 0 or hi == m
        lo_prev = hi
    assert lo_prev == m


def test_expression_source_compiles_without_a_gpu():
    """``ddb_features_source_check``: NVRTC runs on the host, so the expressions of an ``expression_basis`` are
    validated (and the compiler's diagnostics returned) before any GPU is involved."""
    import ctypes as C

    from ddb200 import _lib

    lib = _lib.load()
    fs = ddb200.FeatureSource(["x[0]", "sin(p[0] * x[0]) * x[1]", "exp(-x[0] * x[0])"], params=[0.7])
    log = C.create_string_buffer(4096)
    assert lib.ddb_features_source_check(fs.body.encode(), log, 4096) == 0
    assert lib.ddb_features_source_check(b"f[0] = undefined_symbol;", log, 4096) == -1
    assert b"undefined" in log.value
    b = ddb200.expression_basis(2, fs.expressions, params=[0.7])
    assert len(b) == 3 and b.n_states == 2 and b.term_strings()[1] == "(sin(p[0] * x[0]) * x[1])"


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the base contract's
    keys, `impl`, a `cpu_baseline` that describes this very run and an `e2e` that repeats the line's value with no
    copies; runs without a GPU."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "samples" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0
