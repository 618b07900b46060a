"""CPU checks of the B-row-panel plan (crp-spmm_b200/csrc/cuda/{rowgroup_build,panel_build}.hpp): the same
structures the sm_100a panel kernel consumes are built and decoded on the host by tests/native/panel_emul.cpp
(test infrastructure) and the product they describe is compared BIT FOR BIT with the oracle's CSR loop
(reference local SpMM: src/rowpara_spmm.c:398-408 through the MKL stand-in semantics)."""
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
    so = os.path.join(str(tmp_path_factory.mktemp("native")), "libpanel_emul.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "crp-spmm_b200", "csrc", "cuda"),
                           os.path.join(ROOT, "tests", "native", "panel_emul.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.panel_emul_spmm.restype = C.c_int
    lib.panel_emul_spmm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p]
    return lib


def run(emul, mat, n, K=8, CR=32, EMAX=128, fill=1.0, forced=0, cluster=0):
    m, k, rp, ci, v = mat
    emul.panel_emul_set_cluster(cluster)
    rp, ci = O.i32(rp), O.i32(ci)
    v = np.ascontiguousarray(v, dtype=np.float64)
    B = gen.fill_B(0, k, 0, n)
    Cp = np.full((m, n), np.nan)
    st = np.zeros(11, dtype=np.int64)
    rc = emul.panel_emul_spmm(m, k, O.p(rp), O.p(ci), O.p(v), n, O.p(B), O.p(Cp), K, CR, EMAX, fill, forced, O.p(st))
    assert rc == 0, rc
    Cref = np.zeros((m, n))
    O.lib().orc_csr_spmm(m, n, O.p(rp), O.p(ci), O.p(v), O.p(B), n, O.p(Cref), n)
    assert np.array_equal(Cp, Cref)
    return dict(R=int(st[0]), ngroups=int(st[1]), nblk=int(st[2]), nrest=int(st[3]), ntiles=int(st[4]), nchunks=int(st[5]),
                union_rows=int(st[6]), relaxed=int(st[7]), meta=int(st[8]), max_rows=int(st[9]), max_ent=int(st[10]))


def perturb(mat, frac, seed=5):
    """drop `frac` of the off-diagonal entries: what boundary conditions do to the node blocks of a real FEM matrix"""
    m, k, rp, ci, v = mat
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(m), np.diff(rp))
    keep = (rng.random(ci.size) >= frac) | (rows == ci)
    nrp = np.zeros(m + 1, np.int64)
    nrp[1:] = np.cumsum(np.bincount(rows[keep], minlength=m))
    return m, k, nrp.astype(np.int32), ci[keep], v[keep]


@pytest.mark.parametrize("K,CR,EMAX", [(8, 32, 128), (8, 16, 48), (8, 4, 8), (8, 64, 256)])
def test_pwtk_like_exact_groups(emul, K, CR, EMAX):
    mat = gen.pwtk_like(m=6000, target_nnz=316000, bandwidth=5000, grid_w=16, seed=11)
    s = run(emul, mat, 8, K, CR, EMAX)
    assert s["R"] == 6 and s["nrest"] <= 12 and s["relaxed"] == 0      # the two nodes of the trimmed far coupling are not exact groups
    assert s["max_rows"] <= CR and s["max_ent"] <= EMAX
    # the point of the panel: neighbouring groups share B rows, each is staged once per tile
    assert s["nblk"] / s["union_rows"] > (1.5 if K == 8 else 1.7)


def test_misaligned_first_row_and_rest_rows(emul):
    m, k, rp, ci, v = gen.pwtk_like(m=3000, target_nnz=150000, bandwidth=2500, grid_w=12, seed=3)
    # a rank's slice starts in the middle of a node: rows 4 .. m
    lo = 4
    mat = (m - lo, k, (rp[lo:] - rp[lo]).astype(np.int32), ci[rp[lo]:], v[rp[lo]:])
    s = run(emul, mat, 4)
    assert s["R"] == 6 and 0 < s["nrest"] <= 20


@pytest.mark.parametrize("frac", [0.01, 0.05])
def test_relaxed_groups_on_perturbed_matrix(emul, frac):
    mat = perturb(gen.pwtk_like(m=6000, target_nnz=316000, bandwidth=5000, grid_w=16, seed=11), frac)
    exact = run(emul, mat, 4, fill=1.0)
    relaxed = run(emul, mat, 4, fill=0.75)
    m = mat[0]
    # with exact matching most groups are lost to the row-split kernel; masked blocks keep them
    assert relaxed["R"] == 6 and relaxed["relaxed"] > 0
    assert relaxed["nrest"] < 0.02 * m
    assert exact["R"] == 1 or exact["nrest"] > 5 * max(relaxed["nrest"], 1)


def test_stencil_relaxed_low_fill(emul):
    mat = gen.stencil27(n=12)
    s = run(emul, mat, 4, fill=0.3)
    assert s["R"] > 1 and s["relaxed"] > 0


@pytest.mark.parametrize("name", ["random", "tridiag", "empty_rows"])
def test_general_matrices(emul, name):
    if name == "random":
        mat = gen.random_rect(300, 350, 6, seed=1)
    elif name == "tridiag":
        mat = gen.tridiag(64)
    else:
        mat = gen.random_rect(200, 200, 5, seed=2, empty_rows=(0, 7, 8, 199))
    for fill in (1.0, 0.5):
        run(emul, mat, 3, fill=fill)


def test_forced_group_sizes(emul):
    mat = gen.pwtk_like(m=2400, target_nnz=120000, bandwidth=2000, grid_w=10, seed=7)
    for R in (2, 3, 6):
        s = run(emul, mat, 2, forced=R)
        assert s["R"] == R


@pytest.mark.parametrize("which", ["pwtk", "stencil", "random", "misaligned"])
def test_clustered_tiles(emul, which):
    """Tiles formed by column overlap (crp_panel_cluster_tiles): every group in exactly one tile, the same bits as the CSR loop,
    and - the point - smaller B row panels than tiles of K consecutive groups on mesh-like matrices."""
    if which == "pwtk":
        mat, kw = gen.pwtk_like(m=12000, target_nnz=632000, bandwidth=10000, grid_w=32, seed=11), {}
    elif which == "stencil":
        mat, kw = gen.stencil27(n=16), dict(fill=0.3)
    elif which == "random":
        mat, kw = gen.random_rect(300, 350, 6, seed=1), dict(fill=0.5)
    else:
        m, k, rp, ci, v = gen.pwtk_like(m=3000, target_nnz=150000, bandwidth=2500, grid_w=12, seed=3)
        mat, kw = (m - 4, k, (rp[4:] - rp[4]).astype(np.int32), ci[rp[4]:], v[rp[4]:]), {}
    seq = run(emul, mat, 4, cluster=0, **kw)
    clu = run(emul, mat, 4, cluster=1, **kw)
    assert (clu["R"], clu["ngroups"], clu["nblk"], clu["ntiles"]) == (seq["R"], seq["ngroups"], seq["nblk"], seq["ntiles"])
    if which == "pwtk":
        assert clu["union_rows"] < 0.9 * seq["union_rows"], (clu["union_rows"], seq["union_rows"])
    if which == "stencil":
        assert clu["union_rows"] < 0.75 * seq["union_rows"], (clu["union_rows"], seq["union_rows"])
