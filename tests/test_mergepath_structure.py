"""CPU check of the nnz-balanced partition of spmm_mergepath.cu (mergepath_build.hpp): every row and every nonzero
is covered exactly once, no chunk exceeds ITEMS merge items, only rows longer than a chunk are split, and their
segments are listed in ascending order (the fix-up pass adds them in that order: deterministic, no atomics)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from pycrp import gen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = os.path.join(str(tmp_path_factory.mktemp("native")), "libmergepath_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "crp-spmm_b200", "csrc", "cuda"),
                           os.path.join(ROOT, "tests", "native", "mergepath_emul.cpp"), "-o", so])
    return C.CDLL(so)


def partition(emul, rowptr, items):
    rp = O.i32(rowptr)
    m = rp.size - 1
    cap = 4 * (m + int(rp[-1]) // max(items, 1) + 8)
    desc, lrow, lptr, cnt = np.zeros(cap, np.int32), np.zeros(m + 1, np.int32), np.zeros(m + 2, np.int32), np.zeros(3, np.int32)
    assert emul.mergepath_partition(m, O.p(rp), items, O.p(desc), cap, O.p(lrow), O.p(lptr), O.p(cnt)) == 0
    return desc[:4 * cnt[0]].reshape(-1, 4), lrow[:cnt[1]], lptr[:cnt[1] + 1], int(cnt[2])


@pytest.mark.parametrize("items", [4, 32, 256])
@pytest.mark.parametrize("name", ["rmat", "rand", "empty", "onelong"])
def test_partition_invariants(emul, name, items):
    if name == "rmat":
        rp = gen.rmat(scale=11, edge_factor=16, seed=5)[2]
    elif name == "rand":
        rp = gen.random_rect(700, 500, 9, seed=1, empty_rows=(0, 13, 699))[2]
    elif name == "empty":
        rp = np.zeros(50, np.int32)
    else:
        rp = np.array([0, 0, 5000, 5001, 5001], np.int32)
    m, nnz = rp.size - 1, int(rp[-1])
    desc, lrow, lptr, nseg = partition(emul, rp, items)
    row_seen = np.zeros(m, np.int32)
    nz_seen = np.zeros(nnz, np.int32)
    slots = []
    for r0, y, p0, p1 in desc:
        if y > 0:
            assert p0 == rp[r0] and p1 == rp[r0 + y]
            assert (p1 - p0) + y <= items                      # nonzeros + row ends
            row_seen[r0:r0 + y] += 1
        else:
            assert 0 < p1 - p0 <= items and rp[r0] <= p0 and p1 <= rp[r0 + 1]
            assert rp[r0 + 1] - rp[r0] + 1 > items            # only rows longer than a chunk are split
            slots.append((r0, -y - 1, p0))
        nz_seen[p0:p1] += 1
    assert np.all(nz_seen == 1)
    row_seen[lrow] += 1
    assert np.all(row_seen == 1)
    # scratch slots: 0 .. nseg-1, ascending with the nonzero position inside each long row, ranges as listed
    assert [s for _, s, _ in slots] == list(range(nseg))
    for i, r in enumerate(lrow):
        mine = [(s, p) for rr, s, p in slots if rr == r]
        assert [s for s, _ in mine] == list(range(lptr[i], lptr[i + 1]))
        assert [p for _, p in mine] == sorted(p for _, p in mine)
