"""GPU-side plan construction (crp-spmm_b200/csrc/cuda/plan_build.cu) against the host code it replaces:
  crp_cuda_plan_needed_rows  <- the O(nnz) sweeps of rp_spmm_init (reference src/rowpara_spmm.c:46-112)
  crp_cuda_part_comm_size    <- csr_mat_row_part_comm_size (reference src/spmat_part.c:38-64)
Integers only: everything must be identical - to numpy restatements, to the library's own host path, and, through the whole
engine with the device path forced on (CRP_SPMM_GPU_PLAN_MIN_NNZ=0), to the reference's golden plans and C."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
from pycrp import capi, gen
from util import run_flow

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def lib():
    L = capi.load()
    ip = C.POINTER(C.c_int)
    L.crp_cuda_plan_needed_rows.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, ip, ip, ip, C.POINTER(ip)]
    L.crp_cuda_plan_needed_rows.restype = C.c_int
    L.crp_cuda_part_comm_size.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, ip]
    L.crp_cuda_part_comm_size.restype = C.c_int
    L.csr_mat_row_part_comm_size.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, ip]
    return L


MATS = {
    "pwtk": lambda: gen.pwtk_like(m=12000, target_nnz=632000, bandwidth=10000, grid_w=32, seed=11),
    "er": lambda: gen.erdos_renyi(scale=15, nnz_per_row=16, seed=1),
    "rmat": lambda: gen.rmat(scale=13, edge_factor=16, seed=2),
    "rect": lambda: gen.random_rect(700, 1900, 9, seed=4, empty_rows=(0, 5, 699)),
}


@pytest.mark.parametrize("name", sorted(MATS))
@pytest.mark.parametrize("reidx", [1, 0])
def test_needed_rows_and_reindex(name, reidx):
    L = lib()
    m, k, rp, ci, v = MATS[name]()
    # a rank's block: the middle third of the rows
    r0, r1 = m // 3, 2 * m // 3
    col = np.ascontiguousarray(ci[rp[r0]:rp[r1]], dtype=np.int32)
    out = np.empty_like(col)
    lo, hi, nn = C.c_int(), C.c_int(), C.c_int()
    rows_p = C.POINTER(C.c_int)()
    assert L.crp_cuda_plan_needed_rows(capi.ptr(col), col.size, k, reidx, capi.ptr(out), C.byref(lo), C.byref(hi), C.byref(nn), C.byref(rows_p)) == 1
    uniq = np.unique(col)
    assert (lo.value, hi.value, nn.value) == (int(col.min()), int(col.max()), uniq.size)
    rows = np.ctypeslib.as_array(rows_p, (nn.value,)).copy()
    assert np.array_equal(rows, uniq)
    want = np.searchsorted(uniq, col) if reidx else col - col.min()
    assert np.array_equal(out, want.astype(np.int32))


@pytest.mark.parametrize("name", sorted(MATS))
@pytest.mark.parametrize("nblk", [1, 3, 8])
def test_part_comm_size_matches_host_function(name, nblk):
    L = lib()
    m, k, rp, ci, v = MATS[name]()
    rp32, ci32 = np.ascontiguousarray(rp, dtype=np.int32), np.ascontiguousarray(ci, dtype=np.int32)
    rblk = np.zeros(nblk + 1, np.int32)
    L.csr_mat_row_partition(m, capi.ptr(rp32), nblk, capi.ptr(rblk))
    if nblk == 3:
        rblk[1] = rblk[2]                       # an empty block in the middle
    xd = np.array([(k * b) // nblk for b in range(nblk + 1)], np.int32)
    host, dev = np.zeros(nblk, np.int32), np.zeros(nblk, np.int32)
    th, td = C.c_int(), C.c_int()
    # matrices below CRP_SPMM_GPU_PLAN_MIN_NNZ: the public function runs its host loop
    L.csr_mat_row_part_comm_size(m, k, capi.ptr(rp32), capi.ptr(ci32), nblk, capi.ptr(rblk), capi.ptr(xd), capi.ptr(host), C.byref(th))
    assert L.crp_cuda_part_comm_size(m, k, capi.ptr(rp32), capi.ptr(ci32), nblk, capi.ptr(rblk), capi.ptr(xd), capi.ptr(dev), C.byref(td)) == 1
    L.crp_cuda_part_cache_release()
    # numpy restatement: distinct columns of the block's rows outside its own column range
    for b in range(nblk):
        cols = np.unique(ci32[rp32[rblk[b]]:rp32[rblk[b + 1]]])
        assert int(np.sum((cols < xd[b]) | (cols >= xd[b + 1]))) == int(host[b]) == int(dev[b]), b
    assert th.value == td.value == int(host.sum())


@pytest.mark.parametrize("name", ["rand300_2d_np4_n16", "pwtk600_2d_np8_n64", "rmat8_2d_np8_n16", "rand300_rp_np4_n8_noreidx", "rect200x350_2d_np4_n12", "tridiag16_2d_np8_n16"])
def test_engine_with_device_built_plans_matches_reference_golden(name, tmp_path):
    """the whole engine with the device path forced for every size: grids, every plan array and C as the reference's"""
    case = next(c for c in cases.SPMM_CASES if c[0] == name)
    _, spec, n, mode, nproc, layout, reidx = case
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    csr = os.path.join(str(tmp_path), "a.bin")
    gen.write_csr_bin(csr, int(g["m"]), int(g["k"]), g["csr_rowptr"], g["csr_colidx"], g["csr_val"])
    dumps = run_flow(tmp_path, csr, n, mode, nproc, layout, reidx, device=True, extra_env={"CRP_SPMM_GPU_PLAN_MIN_NNZ": "0", "CRP_SPMM_GPU_PLAN": "1"})
    num = den = 0.0
    for r in range(nproc):
        nrow, ncol = int(g[f"r{r}/C_nrow"][0]), int(g[f"r{r}/C_ncol"][0])
        Cref = g[f"r{r}/C"]
        Cref = Cref.reshape(nrow, ncol) if layout == 0 else Cref.reshape(ncol, nrow).T
        num += float(np.sum((dumps[r]["C"] - Cref) ** 2)); den += float(np.sum(Cref ** 2))
        for key in ("rB_scnts", "rB_rcnts", "rB_sridxs", "rB_rridxs", "A_colidx"):
            assert np.array_equal(dumps[r][key], g[f"r{r}/{key}"]), (r, key)
        assert int(dumps[r]["rB_recv_size"]) == int(g[f"r{r}/rB_recv_size"][0])
        for key in ("rB_nrow", "rB_self_nrow", "rB_self_src_offset", "rB_self_dst_offset"):
            assert int(dumps[r][key]) == int(g[f"r{r}/{key}"][0]), (r, key)
        if mode == "2d":
            assert (int(dumps[r]["pm"]), int(dumps[r]["pn"])) == (int(g[f"r{r}/pm"][0]), int(g[f"r{r}/pn"][0]))
            if r == 0:                                                  # the reference driver computes the cost on rank 0 only
                assert int(dumps[r]["comm_cost"]) == int(g[f"r{r}/comm_cost"][0])
            for key in ("A0_rowptr", "B_rowptr", "AC_rowptr", "BC_colptr"):
                assert np.array_equal(dumps[r][key], g[f"r{r}/{key}"]), (r, key)
    assert np.sqrt(num) <= 1e-12 * np.sqrt(den)
