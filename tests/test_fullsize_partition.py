"""The partitioner / cost model at BASELINE.json's full sizes: library == oracle == compiled reference (oracle/_ref/libref_part.so),
bit for bit, plus the grids SURVEY.md App. A.2 predicts analytically (ER 2^22 x 16, n = 64, 8 ranks -> 2 x 4; stencil 128^3,
n = 1024, 8 ranks -> 4 x 2)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from pycrp import capi, gen
from test_partitioner import REF_PART, lib_part2d, ref_funcs

CASES = {
    "pwtk": (lambda: gen.pwtk_like(), 256, [4, 8], {}),
    "er": (lambda: gen.erdos_renyi(scale=22, nnz_per_row=16, seed=1), 64, [8], {8: (2, 4)}),
    "stencil": (lambda: gen.stencil27(128), 1024, [8], {8: (4, 2)}),
}


@pytest.mark.parametrize("name", list(CASES))
def test_fullsize_grid_and_splits(name):
    make, n, nprocs, expect = CASES[name]
    m, k, rp, ci, v = make()
    del v
    L = capi.load()
    R = ref_funcs()
    for nproc in nprocs:
        rb = np.zeros(nproc + 1, np.int32)
        L.csr_mat_row_partition(m, capi.ptr(rp), nproc, capi.ptr(rb))
        assert np.array_equal(rb, O.row_partition(rp, nproc))
        mine = lib_part2d(L, nproc, m, n, k, rb, rp, ci)
        if nproc in expect:
            assert (mine["pm"], mine["pn"]) == expect[nproc]
        others = [("oracle", O.part2d(nproc, m, n, k, rb, rp, ci))]
        if R is not None:
            others.append(("reference", lib_part2d(R, nproc, m, n, k, rb, rp, ci)))
        for who, ref in others:
            for key in ref:
                assert np.array_equal(np.atleast_1d(mine[key]), np.atleast_1d(ref[key])), (who, key, nproc)
