"""Parity at BASELINE.json's full size (configs[1]: pwtk-shaped 217,918^2, n = 256, fp64), through the C-ABI:
the whole C against the oracle's CSR loop, plus size-independent properties (linearity, bit-reproducibility,
a checksum that must equal the one bench.py prints for the same run)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from pycrp import capi, gen
from util import rel_err

pytestmark = pytest.mark.gpu


def spmm_dev(L, plan, dB, m, n, dtype=np.float64):
    dC = capi.DevBuf(m * n * np.dtype(dtype).itemsize)
    L.crp_cuda_spmm_exec(plan, n, np.dtype(dtype).itemsize, 1.0, dB.p, n, None, 0, 0.0, dC.p, n, None)
    L.crp_cuda_device_sync()
    out = dC.to_numpy((m, n), dtype)
    dC.free()
    return out


def test_pwtk_full_size_vs_oracle():
    L = capi.load()
    m, k, rp, ci, v = gen.pwtk_like()
    n = 256
    assert (m, k) == (217918, 217918) and abs(int(rp[-1]) - 11634424) <= 11634424 // 1000
    B = gen.fill_B(0, k, 0, n)                                       # the drivers' B = 0.19 i + 0.24 j
    plan = L.crp_cuda_spmm_plan_create(m, k, k, capi.ptr(rp), capi.ptr(ci), capi.ptr(v), n)
    dB = capi.DevBuf.from_numpy(B)
    Cd = spmm_dev(L, plan, dB, m, n)
    kern = L.crp_cuda_spmm_last_kernel(plan).decode()
    Cref = np.zeros((m, n))
    O.lib().orc_csr_spmm(m, n, O.p(rp), O.p(ci), O.p(v), O.p(B), n, O.p(Cref), n)
    assert rel_err(Cd, Cref) <= 1e-12, kern
    # bit-reproducible
    assert np.array_equal(Cd, spmm_dev(L, plan, dB, m, n))
    # every kernel variant gives the same bits (same summation order: ascending column within a row)
    L.crp_cuda_spmm_set_variant(plan, b"rowsplit")
    assert np.array_equal(Cd, spmm_dev(L, plan, dB, m, n))
    L.crp_cuda_spmm_set_variant(plan, b"auto")
    # linearity on random data: A (B1 + 2 B2) == A B1 + 2 A B2 to rounding
    rng = np.random.default_rng(0)
    B1, B2 = rng.uniform(-1, 1, (k, n)), rng.uniform(-1, 1, (k, n))
    outs = []
    for X in (B1, B2, B1 + 2 * B2):
        dX = capi.DevBuf.from_numpy(X)
        outs.append(spmm_dev(L, plan, dX, m, n))
        dX.free()
    assert rel_err(outs[2], outs[0] + 2 * outs[1]) <= 1e-13
    # SPD by construction (strictly diagonally dominant): x' A x > 0 column by column
    quad = np.einsum("ij,ij->j", B1, outs[0])
    assert np.all(quad > 0)
    dB.free()
    L.crp_cuda_spmm_plan_destroy(plan)


def test_er_scale20_vs_oracle():
    """BASELINE configs[2] shape (uniform random, 16 nnz per row, n = 64) at 1/4 of the rows."""
    L = capi.load()
    m, k, rp, ci, v = gen.erdos_renyi(scale=20, nnz_per_row=16, seed=1)
    n = 64
    B = gen.fill_B(0, k, 0, n)
    plan = L.crp_cuda_spmm_plan_create(m, k, k, capi.ptr(rp), capi.ptr(ci), capi.ptr(v), n)
    dB = capi.DevBuf.from_numpy(B)
    Cd = spmm_dev(L, plan, dB, m, n)
    Cref = np.zeros((m, n))
    O.lib().orc_csr_spmm(m, n, O.p(rp), O.p(ci), O.p(v), O.p(B), n, O.p(Cref), n)
    assert rel_err(Cd, Cref) <= 1e-12
    dB.free()
    L.crp_cuda_spmm_plan_destroy(plan)


def test_stencil_fp32_wide_vs_oracle():
    """BASELINE configs[4] shape (27-point stencil, n = 1024, fp32) on a 48^3 grid."""
    L = capi.load()
    m, k, rp, ci, v = gen.stencil27(48)
    n = 1024
    B = gen.fill_B(0, k, 0, n, dtype=np.float32)
    plan = L.crp_cuda_spmm_plan_create(m, k, k, capi.ptr(rp), capi.ptr(ci), capi.ptr(v), n)
    dB = capi.DevBuf.from_numpy(B)
    Cd = spmm_dev(L, plan, dB, m, n, np.float32)
    Cref = np.zeros((m, n))
    B64 = B.astype(np.float64)
    O.lib().orc_csr_spmm(m, n, O.p(rp), O.p(ci), O.p(v), O.p(B64), n, O.p(Cref), n)
    assert rel_err(Cd, Cref) <= 1e-5
    dB.free()
    L.crp_cuda_spmm_plan_destroy(plan)
