"""GPU parity tests (run on the B200 box): the CUDA path, called through the C-ABI, against
the oracle and the reference's golden dumps.  Tolerances are BASELINE.json's:
||C_ref - C||_F / ||C_ref||_F <= 1e-12 in fp64, <= 1e-5 in fp32."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import oracle_lib as O
from pycrp import capi, gen
from util import rel_err, run_flow

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL64, TOL32 = 1e-12, 1e-5


def oracle_spmm(m, n, rowptr, colidx, val, B):
    Cref = np.zeros((m, n))
    B = np.ascontiguousarray(B, np.float64)
    O.lib().orc_csr_spmm(m, n, O.p(O.i32(rowptr)), O.p(O.i32(colidx)), O.p(np.ascontiguousarray(val, np.float64)), O.p(B), B.shape[1], O.p(Cref), n)
    return Cref


def device_spmm(m, k, rowptr, colidx, val, B, dtype=np.float64, variant=b"auto", alpha=1.0, beta=0.0, C0=None, x0_rows=None, ld_pad=0):
    """C = alpha A B + beta C0 through crp_cuda_spmm_plan_create / crp_cuda_spmm_exec."""
    L = capi.load()
    n = B.shape[1]
    x0 = k if x0_rows is None else x0_rows
    plan = L.crp_cuda_spmm_plan_create(m, k, x0, capi.ptr(O.i32(rowptr)), capi.ptr(O.i32(colidx)), capi.ptr(np.ascontiguousarray(val, np.float64)), n)
    L.crp_cuda_spmm_set_variant(plan, variant)
    ld = n + ld_pad
    Bp = np.zeros((k, ld), dtype)
    Bp[:, :n] = B
    d0 = capi.DevBuf.from_numpy(Bp[:x0])
    d1 = capi.DevBuf.from_numpy(Bp[x0:]) if x0 < k else None
    Cp = np.full((m, ld), -3.0, dtype)
    if C0 is not None:
        Cp[:, :n] = C0
    dC = capi.DevBuf.from_numpy(Cp)
    L.crp_cuda_spmm_exec(plan, n, np.dtype(dtype).itemsize, alpha, d0.p, ld, d1.p if d1 else None, ld, beta, dC.p, ld, None)
    L.crp_cuda_device_sync()
    out = dC.to_numpy((m, ld), dtype)
    name = L.crp_cuda_spmm_last_kernel(plan).decode()
    L.crp_cuda_spmm_plan_destroy(plan)
    for b in (d0, d1, dC):
        if b:
            b.free()
    assert np.all(out[:, n:] == -3.0), "kernel wrote into the padding of C"
    return out[:, :n], name


def _row_slice(mat, r0, r1):
    """Rows [r0, r1) of a matrix: what a rank whose first row is not aligned with the 6-row node blocks holds."""
    m, k, rp, ci, v = mat
    return (r1 - r0, k, (rp[r0:r1 + 1] - rp[r0]).astype(np.int32), ci[rp[r0]:rp[r1]].copy(), v[rp[r0]:rp[r1]].copy())


def _longrows():
    """A few rows far beyond CRP_LONG_ROW (1024) nonzeros among short and empty ones: the segment + reduce path."""
    rng = np.random.default_rng(9)
    m, k = 60, 6000
    lens = rng.integers(0, 12, m)
    lens[[0, 7, 59]] = [5000, 1025, 3000]
    rows = np.repeat(np.arange(m), lens)
    cols = np.concatenate([np.sort(rng.choice(k, int(c), replace=False)) for c in lens])
    vals = rng.uniform(-1, 1, rows.size)
    return (m, k) + gen.coo_to_csr(m, rows.astype(np.int64), cols.astype(np.int64), vals, sum_duplicates=False)


@pytest.fixture(scope="module")
def mats():
    return {
        "rand": gen.random_rect(700, 500, 9, seed=1, empty_rows=(0, 13, 699)),
        "pwtk": gen.pwtk_like(m=3000, target_nnz=155000, bandwidth=2500, grid_w=16, seed=3),
        "rmat": gen.rmat(scale=11, edge_factor=16, seed=5),
        "stencil": gen.stencil27(10),
        "longrows": _longrows(),
        "pwtk_shift": _row_slice(gen.pwtk_like(m=3000, target_nnz=155000, bandwidth=2500, grid_w=16, seed=3), 2, 2999),
        "onerow": (3, 900) + gen.coo_to_csr(3, np.zeros(900, np.int64), np.arange(900, dtype=np.int64), np.linspace(-1, 1, 900)),
    }


@pytest.mark.parametrize("name", ["rand", "pwtk", "pwtk_shift", "rmat", "stencil", "onerow", "longrows"])
@pytest.mark.parametrize("n", [1, 2, 3, 8, 16, 30, 32, 64, 100, 128, 256, 320])
def test_kernel_fp64_vs_oracle(mats, name, n):
    m, k, rp, ci, v = mats[name]
    rng = np.random.default_rng(n)
    B = rng.uniform(-1, 1, (k, n))
    Cref = oracle_spmm(m, n, rp, ci, v, B)
    Cd, kern = device_spmm(m, k, rp, ci, v, B, ld_pad=(3 if n % 2 else 2))
    assert rel_err(Cd, Cref) <= TOL64, kern
    empty = np.diff(rp) == 0
    assert np.all(Cd[empty] == 0.0)                      # beta = 0: rows without nonzeros are written as zeros


@pytest.mark.parametrize("name", ["rand", "pwtk", "rmat", "longrows"])
@pytest.mark.parametrize("n", [4, 24, 64, 256, 1024])
def test_kernel_fp32_vs_oracle(mats, name, n):
    m, k, rp, ci, v = mats[name]
    rng = np.random.default_rng(n)
    B = rng.uniform(-1, 1, (k, n)).astype(np.float32)
    Cref = oracle_spmm(m, n, rp, ci, v.astype(np.float32).astype(np.float64), B.astype(np.float64))
    Cd, kern = device_spmm(m, k, rp, ci, v, B, dtype=np.float32)
    assert rel_err(Cd, Cref) <= TOL32, kern


@pytest.mark.parametrize("variant", [b"rowsplit", b"rowgroup", b"panel", b"mergepath"])
@pytest.mark.parametrize("name", ["rand", "pwtk", "rmat"])
def test_kernel_variants(mats, name, variant):
    m, k, rp, ci, v = mats[name]
    B = np.random.default_rng(0).uniform(-1, 1, (k, 64))
    Cref = oracle_spmm(m, 64, rp, ci, v, B)
    Cd, kern = device_spmm(m, k, rp, ci, v, B, variant=variant)
    assert rel_err(Cd, Cref) <= TOL64, kern
    # a forced variant must really be the kernel that ran (row-group forms exist only for the block-structured matrix)
    want = {b"rowsplit": "rowsplit", b"mergepath": "mergepath", b"rowgroup": "rowgroup" if name == "pwtk" else "rowsplit",
            b"panel": "panel" if name == "pwtk" else "rowsplit"}[variant]
    assert want in kern, (variant, kern)


@pytest.mark.parametrize("name", ["rand", "longrows", "pwtk"])
def test_alpha_beta_and_two_piece_x(mats, name):
    m, k, rp, ci, v = mats[name]
    rng = np.random.default_rng(2)
    B, C0 = rng.uniform(-1, 1, (k, 48)), rng.uniform(-1, 1, (m, 48))
    Cref = 0.5 * oracle_spmm(m, 48, rp, ci, v, B) - 2.0 * C0
    Cd, _ = device_spmm(m, k, rp, ci, v, B, alpha=0.5, beta=-2.0, C0=C0, x0_rows=123)
    assert rel_err(Cd, Cref) <= TOL64


@pytest.mark.parametrize("name", ["pwtk", "pwtk_shift"])
def test_rowgroup_kernel_is_selected_for_block_structured_rows(mats, name):
    """6-dof node blocks are found whatever the alignment of the first local row."""
    m, k, rp, ci, v = mats[name]
    B = np.random.default_rng(3).uniform(-1, 1, (k, 256))
    Cd, kern = device_spmm(m, k, rp, ci, v, B)
    assert "panel_f64_R6" in kern, kern          # the B-row-panel form of the row groups (n >= 64, 16-byte aligned operands)
    assert rel_err(Cd, oracle_spmm(m, 256, rp, ci, v, B)) <= TOL64


def test_linearity_and_determinism(mats):
    """Size-independent properties: A(B1 + 2 B2) = A B1 + 2 A B2 (to rounding); two runs are bit-identical."""
    m, k, rp, ci, v = mats["pwtk"]
    rng = np.random.default_rng(4)
    B1, B2 = rng.uniform(-1, 1, (k, 128)), rng.uniform(-1, 1, (k, 128))
    C1, _ = device_spmm(m, k, rp, ci, v, B1)
    C2, _ = device_spmm(m, k, rp, ci, v, B2)
    C3, _ = device_spmm(m, k, rp, ci, v, B1 + 2 * B2)
    C3b, _ = device_spmm(m, k, rp, ci, v, B1 + 2 * B2)
    assert np.array_equal(C3, C3b)
    assert rel_err(C3, C1 + 2 * C2) <= 1e-13


def test_data_movement_kernels():
    L = capi.load()
    rng = np.random.default_rng(0)
    for dt in (np.float64, np.float32):
        for nrow, ncol, lds, ldd in ((37, 5, 9, 7), (128, 64, 64, 64), (1000, 33, 40, 35), (1, 1, 1, 1)):
            src = rng.uniform(-1, 1, (nrow, lds)).astype(dt)
            dsrc, ddst = capi.DevBuf.from_numpy(src), capi.DevBuf.from_numpy(np.full((nrow, ldd), 9, dt))
            L.crp_cuda_copy_matrix(dt().itemsize, nrow, ncol, dsrc.p, lds, ddst.p, ldd)
            out = ddst.to_numpy((nrow, ldd), dt)
            assert np.array_equal(out[:, :ncol], src[:, :ncol]) and np.all(out[:, ncol:] == 9)
            idx = rng.integers(0, nrow, 50).astype(np.int32)
            didx, dg = capi.DevBuf.from_numpy(idx), capi.DevBuf.from_numpy(np.zeros((50, ncol), dt))
            L.crp_cuda_gather_rows(dt().itemsize, 50, ncol, dsrc.p, lds, didx.p, dg.p, ncol, None)
            L.crp_cuda_device_sync()
            assert np.array_equal(dg.to_numpy((50, ncol), dt), src[idx, :ncol])
            dT = capi.DevBuf.from_numpy(np.zeros((ncol, nrow), dt))
            L.crp_cuda_transpose(dt().itemsize, nrow, ncol, dsrc.p, lds, dT.p, nrow, None)
            L.crp_cuda_device_sync()
            assert np.array_equal(dT.to_numpy((ncol, nrow), dt), src[:, :ncol].T)
            for b in (dsrc, ddst, didx, dg, dT):
                b.free()


def test_host_in_host_out_proxy_call(mats):
    """crp_cuda_csr_spmm_host keeps the deprecated proxy's argument list (deprecated/src/cuda_proxy.cu:122-182)."""
    m, k, rp, ci, v = mats["rand"]
    B = np.random.default_rng(1).uniform(-1, 1, (k, 20))
    Cout = np.zeros((m, 20))
    capi.load().crp_cuda_csr_spmm_host(m, 20, k, 1.0, int(rp[-1]), capi.ptr(O.i32(rp)), capi.ptr(O.i32(ci)), capi.ptr(v), capi.ptr(B), 20, 0.0, capi.ptr(Cout), 20)
    assert rel_err(Cout, oracle_spmm(m, 20, rp, ci, v, B)) <= TOL64


# ---- the whole engine on P ranks (sharing this box's GPU: host-staged exchange) vs the reference's golden C ----
# host buffers (the reference's calling convention) for every golden case, device-resident buffers for one case per shape / flow
DEV_CASES = {"tridiag16_2d_np8_n64", "rand300_2d_np1_n16", "rand300_2d_np4_n16_cm", "rand300_2d_np6_n24", "rand300_rp_np4_n8_noreidx",
             "rect350x200_2d_np6_n32", "stencil6_2d_np8_n32", "rmat8_rp_np4_n16", "pwtk600_2d_np8_n64", "blockdiag_rp_np4_n8"}
ENGINE_RUNS = [(c, False) for c in cases.SPMM_CASES] + [(c, True) for c in cases.SPMM_CASES if c[0] in DEV_CASES]


@pytest.mark.parametrize("case,device", ENGINE_RUNS, ids=[f"{c[0]}-{'devBC' if d else 'hostBC'}" for c, d in ENGINE_RUNS])
def test_engine_matches_reference_golden(case, device, tmp_path):
    name, spec, n, mode, nproc, layout, reidx = case
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    csr = os.path.join(str(tmp_path), "a.bin")
    gen.write_csr_bin(csr, int(g["m"]), int(g["k"]), g["csr_rowptr"], g["csr_colidx"], g["csr_val"])
    dumps = run_flow(tmp_path, csr, n, mode, nproc, layout, reidx, device=device)
    num = den = 0.0
    for r in range(nproc):
        nrow, ncol = int(g[f"r{r}/C_nrow"][0]), int(g[f"r{r}/C_ncol"][0])
        Cref = g[f"r{r}/C"]
        Cref = Cref.reshape(nrow, ncol) if layout == 0 else Cref.reshape(ncol, nrow).T
        Cm = dumps[r]["C"]
        assert Cm.shape == Cref.shape
        num += float(np.sum((Cm - Cref) ** 2)); den += float(np.sum(Cref ** 2))
        for key in ("rB_scnts", "rB_rcnts", "rB_sridxs", "rB_rridxs", "A_colidx"):     # plan unchanged by having a device
            assert np.array_equal(dumps[r][key], g[f"r{r}/{key}"])
        assert int(dumps[r]["rB_recv_size"]) == int(g[f"r{r}/rB_recv_size"][0])
    assert np.sqrt(num) <= TOL64 * np.sqrt(den)


@pytest.mark.parametrize("nproc,mode", [(1, "2d"), (4, "2d"), (3, "rp")])
def test_engine_fp32(nproc, mode, tmp_path):
    m, k, rp, ci, v = gen.stencil27(8)
    csr = os.path.join(str(tmp_path), "a.bin")
    gen.write_csr_bin(csr, m, k, rp, ci, v)
    dumps = run_flow(tmp_path, csr, 40, mode, nproc, f32=True, device=True)
    sim = O.Simulation(m, k, rp, ci, v.astype(np.float32).astype(np.float64), 40, mode, nproc)
    Cs = sim.exec(dtype=np.float32)
    num = sum(float(np.sum((dumps[r]["C"].astype(np.float64) - Cs[r]) ** 2)) for r in range(nproc))
    den = sum(float(np.sum(Cs[r] ** 2)) for r in range(nproc))
    assert np.sqrt(num) <= TOL32 * np.sqrt(den)
    sim.close()
