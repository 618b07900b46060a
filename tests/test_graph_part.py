"""The native k-way partitioner behind METIS's two entry points (crp-spmm_b200/csrc/ingest/graph_part.c, libcrpingest.so) - what the
reference's driver front-end calls for <part-method> = 1 (examples/metis_mat_part.c:31-113).  METIS is not available, so there is
no parity to check: the tests pin what a caller relies on - a valid partition within the imbalance bound, determinism, a cut far
below a random one on meshes whose numbering was shuffled - and run the reference's own METIS_row_partition (compiled from
/root/reference where that tree exists) on top of it: a symmetric permutation P A P' with contiguous balanced row blocks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from pycrp import capi, gen
from util import PKG

REF = "/root/reference/examples"
I32 = np.int32


@pytest.fixture(scope="module")
def lib():
    L = C.CDLL(os.path.join(PKG, "lib", "libcrpingest.so"))
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
    L.METIS_SetDefaultOptions.argtypes = [ip]
    L.METIS_PartGraphKway.argtypes = [ip, ip, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, ip, C.c_void_p, fp, ip, ip, C.c_void_p]
    L.METIS_PartGraphKway.restype = C.c_int
    return L


def partition(L, A, k, ub=1.05, vol=True):
    A = sp.csr_matrix(A)
    n = A.shape[0]
    xadj, adj = np.ascontiguousarray(A.indptr, dtype=I32), np.ascontiguousarray(A.indices, dtype=I32)
    opts = (C.c_int * 40)()
    L.METIS_SetDefaultOptions(opts)
    opts[1] = 1 if vol else 0
    part = np.full(n, -7, dtype=I32)
    nv, nc, kk, obj, u = C.c_int(n), C.c_int(1), C.c_int(k), C.c_int(-1), C.c_float(ub)
    rc = L.METIS_PartGraphKway(C.byref(nv), C.byref(nc), capi.ptr(xadj), capi.ptr(adj), None, None, None, C.byref(kk), None, C.byref(u), opts, C.byref(obj), capi.ptr(part))
    assert rc == 1
    return part, obj.value


def grid_graph(w, h, shuffle_seed=None):
    idx = np.arange(w * h).reshape(h, w)
    r = np.concatenate([idx[:, :-1].ravel(), idx[:-1, :].ravel()])
    c = np.concatenate([idx[:, 1:].ravel(), idx[1:, :].ravel()])
    if shuffle_seed is not None:
        p = np.random.default_rng(shuffle_seed).permutation(w * h)
        r, c = p[r], p[c]
    A = sp.coo_matrix((np.ones(r.size), (r, c)), shape=(w * h, w * h))
    return sp.csr_matrix(A + A.T)


def edge_cut(A, part):
    A = sp.coo_matrix(A)
    return int(np.sum((part[A.row] != part[A.col]) & (A.row < A.col)))


def comm_volume(A, part):
    A = sp.coo_matrix(A)
    off = part[A.row] != part[A.col]
    return np.unique(np.stack([A.row[off], part[A.col[off]]]), axis=1).shape[1]


@pytest.mark.parametrize("k", [1, 2, 4, 7])
def test_valid_balanced_deterministic(lib, k):
    A = grid_graph(40, 30, shuffle_seed=3)
    n = A.shape[0]
    part, obj = partition(lib, A, k)
    assert part.min() >= 0 and part.max() < k and len(np.unique(part)) == k
    sizes = np.bincount(part, minlength=k)
    assert sizes.max() <= 1.05 * n / k + 1, sizes
    assert obj == comm_volume(A, part)                                   # METIS_OBJTYPE_VOL: the reported objective is the real one
    part2, obj2 = partition(lib, A, k)
    assert np.array_equal(part, part2) and obj == obj2
    _, cut = partition(lib, A, k, vol=False)
    assert cut == edge_cut(A, part)


def test_quality_on_a_shuffled_mesh(lib):
    """numbering destroyed, structure intact: the BFS order finds the mesh again - the cut is a few mesh lines, not a random one"""
    A = grid_graph(32, 32, shuffle_seed=11)
    part, _ = partition(lib, A, 4)
    cut = edge_cut(A, part)
    rnd = edge_cut(A, np.random.default_rng(0).integers(0, 4, A.shape[0]).astype(I32))
    assert cut <= 5 * 32, cut                  # three straight cuts through a 32 x 32 mesh would be 96 edges
    assert cut * 6 < rnd, (cut, rnd)


def test_components_isolated_vertices_and_more_parts_than_vertices(lib):
    A = sp.block_diag([grid_graph(6, 5), grid_graph(4, 4), sp.csr_matrix((3, 3))]).tocsr()      # two meshes + three isolated vertices
    part, _ = partition(lib, A, 3)
    assert part.min() >= 0 and part.max() < 3
    assert np.bincount(part, minlength=3).max() <= 1.05 * A.shape[0] / 3 + 1
    tiny = grid_graph(2, 2)
    part, _ = partition(lib, tiny, 8)
    assert part.min() >= 0 and part.max() < 8


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "metis_mat_part.c")), reason="needs /root/reference (build container)")
def test_reference_front_end_on_top_of_it(lib, tmp_path):
    """examples/metis_mat_part.c (unchanged) + libcrpingest: perm is a permutation, row blocks contiguous and balanced, the matrix
    comes back as P A P' with sorted rows, and on a shuffled FEM pattern the communication volume of the METIS-style row blocks is
    a fraction of what the natural split of the shuffled matrix needs."""
    so = str(tmp_path / "libfront.so")
    ing = os.path.join(PKG, "csrc", "ingest")
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-I" + ing, "-I" + REF, "-I/root/reference/src", "-I" + os.path.join(PKG, "..", "include"),
                           os.path.join(REF, "metis_mat_part.c"), "-o", so, "-L" + os.path.join(PKG, "lib"), "-lcrpingest", "-Wl,-rpath," + os.path.join(PKG, "lib")])
    F = C.CDLL(so)
    F.METIS_row_partition.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    m, k, rp, ci, v = gen.pwtk_like(m=6000, target_nnz=316000, bandwidth=5000, grid_w=16, seed=11)
    A = sp.csr_matrix((v, ci, rp), shape=(m, m))
    shuf = np.random.default_rng(5).permutation(m)
    S = sp.csr_matrix(A[shuf][:, shuf])
    S.sort_indices()
    nproc = 4
    rp2, ci2, v2 = S.indptr.astype(I32).copy(), S.indices.astype(I32).copy(), S.data.astype(np.float64).copy()
    perm, displs = np.zeros(m, I32), np.zeros(nproc + 1, I32)
    F.METIS_row_partition(m, nproc, capi.ptr(rp2), capi.ptr(ci2), capi.ptr(v2), capi.ptr(perm), capi.ptr(displs))
    assert np.array_equal(np.sort(perm), np.arange(m))
    assert displs[0] == 0 and displs[-1] == m and np.all(np.diff(displs) <= 1.05 * m / nproc + 1)
    P = sp.csr_matrix((np.ones(m), (perm, np.arange(m))), shape=(m, m))          # row i -> row perm[i]
    want = sp.csr_matrix(P @ S @ P.T)
    want.sort_indices()
    got = sp.csr_matrix((v2, ci2, rp2), shape=(m, m))
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices) and np.allclose(got.data, want.data, rtol=0, atol=0)
    # communication volume (columns outside a block's own range, the library's own counter) before / after
    L = capi.load()
    L.csr_mat_row_part_comm_size.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    def volume(indptr, indices, blocks):
        sizes, tot = np.zeros(nproc, I32), C.c_int()
        L.csr_mat_row_part_comm_size(m, m, capi.ptr(np.ascontiguousarray(indptr, I32)), capi.ptr(np.ascontiguousarray(indices, I32)), nproc,
                                     capi.ptr(np.ascontiguousarray(blocks, I32)), capi.ptr(np.ascontiguousarray(blocks, I32)), capi.ptr(sizes), C.byref(tot))
        return tot.value
    nat = np.zeros(nproc + 1, I32)
    L.csr_mat_row_partition.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.csr_mat_row_partition(m, capi.ptr(S.indptr.astype(I32)), nproc, capi.ptr(nat))
    before, after = volume(S.indptr, S.indices, nat), volume(rp2, ci2, displs)
    assert after * 3 < before, (before, after)
