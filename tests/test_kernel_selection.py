"""Plan-time kernel selection (spmm_rowgroup.cu / spmm_longrow.cu) through its host-only C-ABI view crp_cuda_spmm_analyse:
no device needed.  Pins the decisions DESIGN.md describes: 6-row groups for the pwtk-shaped matrix whatever the alignment of the
first local row, row-split only for unstructured / stencil matrices, segment splitting for power-law hubs."""
import ctypes as C

import numpy as np
import pytest

import cases
from pycrp import capi, gen


def analyse(rp, ci):
    L = capi.load()
    rp, ci = np.ascontiguousarray(rp, np.int32), np.ascontiguousarray(ci, np.int32)
    R, off, nlong, nblk = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
    L.crp_cuda_spmm_analyse(len(rp) - 1, capi.ptr(rp), capi.ptr(ci), C.byref(R), C.byref(off), C.byref(nblk), C.byref(nlong))
    return R.value, off.value, nblk.value, nlong.value


def row_slice(rp, ci, r0, r1):
    return (rp[r0:r1 + 1] - rp[r0]).astype(np.int32), ci[rp[r0]:rp[r1]]


def test_pwtk_like_gets_six_row_groups_at_any_alignment():
    m, k, rp, ci, v = gen.pwtk_like(m=3000, target_nnz=155000, bandwidth=2500, grid_w=16, seed=3)
    R, off, nblk, nlong = analyse(rp, ci)
    assert (R, off, nlong) == (6, 0, 0)
    assert nblk * 6 >= 0.98 * int(rp[-1])                       # nearly every nonzero sits in a 6 x 1 block
    for first in (1, 2, 3, 4, 5, 6, 7):
        srp, sci = row_slice(rp, ci, first, m - 5)
        R, off, nblk, _ = analyse(srp, sci)
        assert R == 6 and off == (6 - first % 6) % 6, (first, R, off)
        assert nblk * 6 >= 0.97 * int(srp[-1])


@pytest.mark.parametrize("spec", [("rand", 700, 500, 9, 1, ()), ("stencil27", 10), ("tridiag", 64)])
def test_unstructured_matrices_keep_the_rowsplit_kernel(spec):
    m, k, rp, ci, v = cases.build_matrix(spec)
    R, off, nblk, nlong = analyse(rp, ci)
    assert (R, nblk, nlong) == (1, 0, 0)


def test_dense_diagonal_blocks_pick_the_block_height():
    m, k, rp, ci, v = cases.build_matrix(("blockdiag", 5, 8))
    assert analyse(rp, ci)[:3] == (8, 0, 5 * 8)                 # 5 groups x 8 columns
    srp, sci = row_slice(rp, ci, 3, m)                          # first local row inside a block
    R, off, nblk, _ = analyse(srp, sci)
    assert (R, off, nblk) == (8, 5, 4 * 8)


def test_power_law_hubs_are_flagged_for_segmenting():
    m, k, rp, ci, v = gen.rmat(scale=13, edge_factor=16, seed=5)
    R, off, nblk, nlong = analyse(rp, ci)
    assert R == 1 and nlong == int(np.sum(np.diff(rp) > 1024)) and nlong > 0


def test_forced_group_size(monkeypatch):
    m, k, rp, ci, v = gen.pwtk_like(m=1200, target_nnz=62000, bandwidth=1000, grid_w=8, seed=7)
    monkeypatch.setenv("CRP_SPMM_ROWGROUP_R", "3")
    assert analyse(rp, ci)[0] == 3
    monkeypatch.setenv("CRP_SPMM_ROWGROUP_R", "1")
    assert analyse(rp, ci)[0] == 1


def test_ranks_per_node_follows_the_launcher(monkeypatch):
    """Transport selection compares the ranks on THIS node (not the world size) with the visible GPUs, so that multi-node jobs
    with one rank per GPU keep the device transports (crp_ranks_on_this_node, csrc/host/crp_common.c)."""
    import ctypes as C
    L = capi.load()
    L.crp_ranks_on_this_node.argtypes = [C.c_int]
    L.crp_ranks_on_this_node.restype = C.c_int
    for var in ("OMPI_COMM_WORLD_LOCAL_SIZE", "MV2_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS", "SLURM_NTASKS_PER_NODE", "LOCAL_WORLD_SIZE"):
        monkeypatch.delenv(var, raising=False)
    assert L.crp_ranks_on_this_node(16) == 16            # nothing known: single node
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert L.crp_ranks_on_this_node(16) == 8
    monkeypatch.setenv("OMPI_COMM_WORLD_LOCAL_SIZE", "4")
    assert L.crp_ranks_on_this_node(16) == 4
    monkeypatch.setenv("OMPI_COMM_WORLD_LOCAL_SIZE", "64")     # nonsense (more than the world): ignored
    assert L.crp_ranks_on_this_node(16) == 8
