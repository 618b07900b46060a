"""spmat_part.c / utils.c of the library vs the oracle (and vs the compiled reference, when
oracle/_ref/libref_part.so is present) on generated CSR patterns: bit-exact integers."""
import ctypes as C
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle_lib as O
from pycrp import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PART = os.path.join(ROOT, "oracle", "_ref", "libref_part.so")


def lib_part2d(L, nproc, m, n, k, rb, rowptr, colidx):
    pm, pn, cost = C.c_int(), C.c_int(), C.c_size_t()
    a0, br, ac, bc = capi.c_int_p(), capi.c_int_p(), capi.c_int_p(), capi.c_int_p()
    L.calc_spmm_part2d_from_1d(nproc, m, n, k, capi.ptr(rb), capi.ptr(rowptr), capi.ptr(colidx), 1, C.byref(pm), C.byref(pn), C.byref(cost),
                               C.byref(a0), C.byref(br), C.byref(ac), C.byref(bc), 0)
    out = dict(pm=pm.value, pn=pn.value, comm_cost=cost.value, A0_rowptr=capi.np_from(a0, nproc + 1, np.int32),
               B_rowptr=capi.np_from(br, pm.value + 1, np.int32), AC_rowptr=capi.np_from(ac, pm.value + 1, np.int32),
               BC_colptr=capi.np_from(bc, pn.value + 1, np.int32))
    return out


def random_csr(rng, m, k, density, empty_frac):
    counts = rng.binomial(k, density, size=m)
    counts[rng.random(m) < empty_frac] = 0
    rowptr = np.zeros(m + 1, np.int32)
    rowptr[1:] = np.cumsum(counts)
    colidx = np.concatenate([np.sort(rng.choice(k, c, replace=False)) for c in counts] + [np.zeros(0, np.int64)]).astype(np.int32)
    return rowptr, colidx


def ref_funcs():
    if not os.path.exists(REF_PART):
        return None
    R = C.CDLL(REF_PART)
    R.calc_spmm_part2d_from_1d.argtypes = capi.load().calc_spmm_part2d_from_1d.argtypes
    R.csr_mat_row_partition.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    return R


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 10 ** 6), m=st.integers(1, 120), nproc=st.sampled_from([1, 2, 3, 4, 6, 8, 12]),
       n=st.sampled_from([1, 4, 16, 64, 256]), square=st.booleans(), empty=st.sampled_from([0.0, 0.2, 0.6]))
def test_partitioner_bit_exact(seed, m, nproc, n, square, empty):
    rng = np.random.default_rng(seed)
    k = m if square else int(rng.integers(1, 150))
    rowptr, colidx = random_csr(rng, m, k, float(rng.uniform(0.02, 0.3)), empty)
    if rowptr[-1] == 0:
        return
    L = capi.load()
    rb = np.zeros(nproc + 1, np.int32)
    L.csr_mat_row_partition(m, capi.ptr(rowptr), nproc, capi.ptr(rb))
    assert np.array_equal(rb, O.row_partition(rowptr, nproc))
    mine = lib_part2d(L, nproc, m, n, k, rb, rowptr, colidx)
    orc = O.part2d(nproc, m, n, k, rb, rowptr, colidx)
    for key in orc:
        assert np.array_equal(np.atleast_1d(mine[key]), np.atleast_1d(orc[key])), key
    R = ref_funcs()
    if R is not None:
        rb2 = np.zeros(nproc + 1, np.int32)
        R.csr_mat_row_partition(m, capi.ptr(rowptr), nproc, capi.ptr(rb2))
        assert np.array_equal(rb, rb2)
        ref = lib_part2d(R, nproc, m, n, k, rb, rowptr, colidx)
        for key in ref:
            assert np.array_equal(np.atleast_1d(mine[key]), np.atleast_1d(ref[key])), ("vs reference", key)


@pytest.mark.parametrize("length,nblk", [(10, 3), (7, 7), (5, 8), (0, 4), (2 ** 31 - 1, 7), (217918, 8)])
def test_calc_block_spos_size(length, nblk):
    L = capi.load()
    sp, sz = C.c_int(), C.c_int()
    got = []
    for i in range(-1, nblk + 2):
        L.calc_block_spos_size(length, nblk, i, C.byref(sp), C.byref(sz))
        got.append((sp.value, sz.value))
        a, b = C.c_int(), C.c_int()
        O.lib().orc_block_spos_size(length, nblk, i, C.byref(a), C.byref(b))
        assert (a.value, b.value) == got[-1]
    assert got[0] == (-1, 0) and got[-1] == (-1, 0) and got[nblk + 1][0] == length


def test_comm_size_and_factorization():
    L = capi.load()
    rng = np.random.default_rng(5)
    rowptr, colidx = random_csr(rng, 90, 70, 0.1, 0.1)
    rb = O.row_partition(rowptr, 6)
    xd = O.block_split(70, 6)
    sizes = np.zeros(6, np.int32)
    tot = C.c_int()
    L.csr_mat_row_part_comm_size(90, 70, capi.ptr(rowptr), capi.ptr(colidx), 6, capi.ptr(rb), capi.ptr(xd), capi.ptr(sizes), C.byref(tot))
    osz, otot = O.comm_size(90, 70, rowptr, colidx, rb, xd)
    assert np.array_equal(sizes, osz) and tot.value == otot
    for n_, exp in ((8, [2, 2, 2]), (12, [2, 2, 3]), (7, [7]), (1, []), (360, [2, 2, 2, 3, 3, 5])):
        f = capi.c_int_p()
        cnt = L.prime_factorization(n_, C.byref(f))
        assert capi.np_from(f, cnt, np.int32).tolist() == exp
