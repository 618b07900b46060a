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


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[2..4] at their FULL sizes: C on 4096 sampled rows against the plain fp64 CSR product of those rows
# (the same check bench.py prints as `parity` for every --workload), plus whole-matrix properties that need no reference.

def sampled_rel_err(Cd, rows, rp, ci, v, n, dtype):
    import scipy.sparse as sp
    lens = (rp[rows + 1] - rp[rows]).astype(np.int64)
    idx = np.concatenate([np.arange(rp[r], rp[r + 1], dtype=np.int64) for r in rows])
    cols, vals = ci[idx].astype(np.int64), v[idx].astype(np.float64)
    if dtype == np.float32:
        vals = vals.astype(np.float32).astype(np.float64)
    ucols, inv = np.unique(cols, return_inverse=True)
    Bu = (ucols.astype(np.float64)[:, None] * 0.19 + np.arange(n, dtype=np.float64)[None, :] * 0.24).astype(dtype).astype(np.float64)
    indptr = np.zeros(rows.size + 1, np.int64)
    indptr[1:] = np.cumsum(lens)
    Cref = sp.csr_matrix((vals, inv, indptr), shape=(rows.size, ucols.size)) @ Bu
    return rel_err(Cd[rows], Cref)


def full_size_case(workload, tol, expect_kernel=None):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    L = capi.load()
    gname, gkw, n, dts, mode, desc = bench.WORKLOADS[workload]
    dt = np.float32 if dts == "f32" else np.float64
    m, k, rp, ci, v = gen.read_csr_bin(bench.matrix_path(workload))          # generated once per box, shared with bench.py
    rp, ci, v = np.ascontiguousarray(rp), np.ascontiguousarray(ci), np.ascontiguousarray(v)
    plan = L.crp_cuda_spmm_plan_create(m, k, k, capi.ptr(rp), capi.ptr(ci), capi.ptr(v), n)
    dB = capi.DevBuf.from_numpy(gen.fill_B(0, k, 0, n, dtype=dt))
    Cd = spmm_dev(L, plan, dB, m, n, dt)
    kern = L.crp_cuda_spmm_last_kernel(plan).decode()
    if expect_kernel:
        assert expect_kernel in kern, kern
    rows = np.sort(np.random.default_rng(7).choice(m, size=4096, replace=False))
    assert sampled_rel_err(Cd, rows, rp, ci, v, n, dt) <= tol, kern
    # column j of B is 0.19 i + 0.24 j: C[:, j] - C[:, 0] = 0.24 j * rowsum(A) for every row - a whole-matrix check without a reference
    rowsum = np.add.reduceat(v, rp[:-1].astype(np.int64)) * (rp[1:] > rp[:-1])
    j = n - 1
    lhs = Cd[:, j].astype(np.float64) - Cd[:, 0].astype(np.float64)
    scale = np.abs(Cd[:, j]).astype(np.float64).max()
    assert np.max(np.abs(lhs - 0.24 * j * rowsum)) <= (1e-9 if dt == np.float64 else 2e-3) * scale, kern
    # a second exec gives the same bits
    assert np.array_equal(Cd, spmm_dev(L, plan, dB, m, n, dt))
    dB.free()
    L.crp_cuda_spmm_plan_destroy(plan)
    return kern


def test_er_scale22_full_size_sampled():
    """BASELINE configs[2]: Erdos-Renyi 4M x 4M, 16 nnz/row, n = 64, fp64."""
    full_size_case("er", 1e-12)


def test_stencil128_fp32_full_size_sampled():
    """BASELINE configs[4]: 27-point stencil 128^3, n = 1024, fp32 (8.6 GB each for B and C)."""
    kern = full_size_case("stencil", 1e-5)
    assert "panel" in kern, kern         # neighbouring stencil rows share 2/3 of their columns: masked row groups, B panels in shared memory


@pytest.mark.skipif(os.environ.get("CRP_TEST_RMAT22", "0") != "1", reason="RMAT-22 takes minutes to generate; CRP_TEST_RMAT22=1 (bench.py --workload rmat prints the same parity)")
def test_rmat_scale22_full_size_sampled():
    """BASELINE configs[3]: RMAT scale 22, edge factor 32, n = 128, fp64 - the nnz-balanced kernel."""
    full_size_case("rmat", 1e-12, expect_kernel="mergepath")


def test_rmat_scale19_sampled_mergepath():
    """The same shape at scale 19 (generated in seconds): the automatic choice must be the nnz-balanced kernel."""
    L = capi.load()
    m, k, rp, ci, v = gen.rmat(scale=19, edge_factor=32)
    n = 128
    plan = L.crp_cuda_spmm_plan_create(m, k, k, capi.ptr(rp), capi.ptr(ci), capi.ptr(v), n)
    dB = capi.DevBuf.from_numpy(gen.fill_B(0, k, 0, n))
    Cd = spmm_dev(L, plan, dB, m, n)
    kern = L.crp_cuda_spmm_last_kernel(plan).decode()
    assert "mergepath" in kern, kern
    rows = np.sort(np.random.default_rng(7).choice(m, size=4096, replace=False))
    rows = np.union1d(rows, np.argsort(np.diff(rp))[-8:])            # the eight longest rows (the hubs) as well
    assert sampled_rel_err(Cd, rows, rp, ci, v, n, np.float64) <= 1e-12
    dB.free()
    L.crp_cuda_spmm_plan_destroy(plan)
