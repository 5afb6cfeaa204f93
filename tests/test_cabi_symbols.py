"""The C-ABI library loads, exports every symbol include/ddb.h declares, and refuses to compute
without a GPU (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ddb200
from ddb200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ddb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddb_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libddb200.so does not export {n}"
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes and the header are out of sync"


def test_version_string():
    assert b"sm_100a" in _lib.load().ddb_version()


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(_lib.DDBError) as ei:
        ddb200.Context(0)
    assert ei.value.status == -3 and "no CPU fallback" in str(ei.value)


def test_struct_layouts():
    assert C.sizeof(_lib.SweepParams) == 48
    assert C.sizeof(_lib.Timings) == 104


@pytest.mark.parametrize("kaug,m,nctas", [(59, 10_000_000, 444), (2209, 1_250_000, 148), (2209, 1001, 148), (128, 777, 148),
                                          (134, 100_000, 148), (300, 16, 148)])
def test_gram_plan_covers_every_stage_exactly_once(kaug, m, nctas):
    lib = _lib.load()
    info = np.zeros(6, dtype=np.int64)
    n = lib.ddb_gram_plan_describe(kaug, m, nctas, None, 0, _lib.ptr(info))
    assert n > 0
    seg = np.zeros((n, 6), dtype=np.int64)
    assert lib.ddb_gram_plan_describe(kaug, m, nctas, _lib.ptr(seg), n, None) == n
    bt, ks, nblk, npairs, nstages, nslots = info.tolist()
    assert nblk == -(-kaug // bt) and npairs == nblk * (nblk + 1) // 2 and nstages == -(-m // ks)
    cover = {}
    for cta, bi, bj, slot, b, e in seg.tolist():
        assert 0 <= cta < nctas and 0 <= bj <= bi < nblk and 0 <= b < e <= nstages
        cover.setdefault((bi, bj), []).append((b, e))
    assert len(cover) == npairs
    for ranges in cover.values():
        ranges.sort()
        assert ranges[0][0] == 0 and ranges[-1][1] == nstages
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(len(ranges) - 1))
    assert sorted(seg[:, 3].tolist()) == list(range(nslots))  # one partial-tile slot per segment
    # balance: no CTA carries more than ~6 % above the mean weighted load
    cost = np.where(seg[:, 1] == seg[:, 2], 0.6 if bt == 128 else 0.85, 1.0)
    last_rows = kaug - (nblk - 1) * bt
    if bt == 128 and last_rows <= 64:
        cost = np.where(seg[:, 1] == nblk - 1, 0.53, cost)
    load = np.bincount(seg[:, 0], weights=cost * (seg[:, 5] - seg[:, 4]), minlength=nctas)
    if nstages * npairs > 50 * nctas:
        assert load.max() <= 1.06 * load.sum() / nctas + 1

