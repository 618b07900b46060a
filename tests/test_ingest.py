"""Fast ingest for the drivers (crp-spmm_b200/csrc/ingest/mmio_fast.c -> libcrpingest.so): drop-in mm_read_sparse_RPI and
coo2csr.  Checked against the reference's own reader (examples/mmio_utils.c:11-190, compiled from /root/reference into a
scratch .so where that tree exists) entry for entry, and against scipy everywhere else: Matrix Market real / integer /
pattern, general / symmetric, comments and blank lines, unsorted input, the binary CSR short-cut."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from pycrp import gen
from util import PKG, ROOT

REF = "/root/reference/examples"


def bind(lib):
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    lib.mm_read_sparse_RPI.argtypes = [C.c_char_p, C.c_int, ip, ip, ip, C.POINTER(ip), C.POINTER(ip), C.POINTER(dp)]
    lib.mm_read_sparse_RPI.restype = C.c_int
    lib.coo2csr.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, dp, C.POINTER(ip), C.POINTER(ip), C.POINTER(dp)]
    return lib


@pytest.fixture(scope="module")
def fast():
    return bind(C.CDLL(os.path.join(PKG, "lib", "libcrpingest.so")))


@pytest.fixture(scope="module")
def reference(tmp_path_factory):
    if not os.path.exists(os.path.join(REF, "mmio_utils.c")):
        return None
    so = str(tmp_path_factory.mktemp("refio") / "librefio.so")
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-I" + REF, "-I/root/reference/src", "-Wno-unused-result",
                           os.path.join(REF, "mmio.c"), os.path.join(REF, "mmio_utils.c"), "-o", so])
    return bind(C.CDLL(so))


def read(lib, path, need_symm=0):
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    m, k, nnz = C.c_int(), C.c_int(), C.c_int()
    r, c, v = ip(), ip(), dp()
    rc = lib.mm_read_sparse_RPI(path.encode(), need_symm, C.byref(m), C.byref(k), C.byref(nnz), C.byref(r), C.byref(c), C.byref(v))
    if rc != 0:
        return rc, None
    n = nnz.value
    coo = (m.value, k.value, np.ctypeslib.as_array(r, (n,)).copy(), np.ctypeslib.as_array(c, (n,)).copy(), np.ctypeslib.as_array(v, (n,)).copy())
    rp, ci, cv = ip(), ip(), dp()
    lib.coo2csr(m.value, k.value, n, r, c, v, C.byref(rp), C.byref(ci), C.byref(cv))
    csr = (np.ctypeslib.as_array(rp, (m.value + 1,)).copy(), np.ctypeslib.as_array(ci, (max(n, 1),))[:n].copy(), np.ctypeslib.as_array(cv, (max(n, 1),))[:n].copy())
    return 0, (coo, csr)


def write_mtx(path, kind, symm, m, k, rows, cols, vals, shuffle_seed=None, comments=True):
    idx = np.arange(rows.size)
    if shuffle_seed is not None:
        np.random.default_rng(shuffle_seed).shuffle(idx)
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {kind} {'symmetric' if symm else 'general'}\n")
        if comments:
            f.write("% a comment line\n%another\n\n")
        f.write(f"{m} {k} {rows.size}\n")
        for i in idx:
            if kind == "pattern":
                f.write(f"{rows[i] + 1} {cols[i] + 1}\n")
            elif kind == "integer":
                f.write(f"{rows[i] + 1}  {cols[i] + 1} {int(vals[i])}\n")
            else:
                f.write(f"{rows[i] + 1} {cols[i] + 1} {vals[i]:.17g}\n")
        if comments:
            f.write("\n")


CASES = [("real", False), ("real", True), ("integer", False), ("pattern", True), ("pattern", False)]


@pytest.mark.parametrize("kind,symm", CASES)
@pytest.mark.parametrize("threads", ["1", "7"])
def test_matrix_market_matches_reference_and_scipy(fast, reference, kind, symm, threads, tmp_path, monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", threads)
    rng = np.random.default_rng(5)
    m = k = 211
    A = sp.random(m, k, density=0.04, random_state=3, format="coo")
    rows, cols = A.row.astype(np.int64), A.col.astype(np.int64)
    if symm:
        keep = rows >= cols
        rows, cols = rows[keep], cols[keep]
    vals = rng.integers(-9, 10, rows.size).astype(np.float64) if kind == "integer" else rng.standard_normal(rows.size) * 10.0 ** rng.integers(-8, 8, rows.size)
    if kind == "pattern":
        vals = np.ones(rows.size)
    path = str(tmp_path / "a.mtx")
    write_mtx(path, kind, symm, m, k, rows, cols, vals, shuffle_seed=11)
    rc, out = read(fast, path, need_symm=1 if symm else 0)
    assert rc == 0
    (mm, kk, r, c, v), (rp, ci, cv) = out
    assert (mm, kk) == (m, k)
    # against scipy
    full_r, full_c, full_v = rows, cols, vals
    if symm:
        off = rows != cols
        full_r, full_c, full_v = np.concatenate([rows, cols[off]]), np.concatenate([cols, rows[off]]), np.concatenate([vals, vals[off]])
    S = sp.csr_matrix((full_v, (full_r, full_c)), shape=(m, k))
    S.sort_indices()
    assert np.array_equal(rp, S.indptr) and np.array_equal(ci, S.indices) and np.array_equal(cv, S.data)
    # against the reference's reader: the COO arrays entry for entry (file order, mirrored entries appended), and the CSR
    if reference is not None:
        rc2, ref = read(reference, path, need_symm=1 if symm else 0)
        assert rc2 == 0
        for a, b in zip(out[0], ref[0]):
            assert np.array_equal(a, b)
        for a, b in zip(out[1], ref[1]):
            assert np.array_equal(a, b)


def test_binary_csr_shortcut_and_sorted_fast_path(fast, tmp_path):
    m, k, rp, ci, v = gen.pwtk_like(m=3000, target_nnz=150000, bandwidth=2500, grid_w=12, seed=3)
    path = str(tmp_path / "a.bin")
    gen.write_csr_bin(path, m, k, rp, ci, v)
    rc, out = read(fast, path)
    assert rc == 0
    (mm, kk, r, c, vv), (rp2, ci2, cv2) = out
    assert (mm, kk) == (m, k)
    assert np.array_equal(rp2, rp) and np.array_equal(ci2, ci) and np.array_equal(cv2, v)
    assert np.array_equal(r, np.repeat(np.arange(m), np.diff(rp)))
    # the same matrix as text
    mtx = str(tmp_path / "a.mtx")
    gen.write_mtx(mtx, m, k, rp, ci, v)
    rc, out = read(fast, mtx)
    assert rc == 0 and np.array_equal(out[1][0], rp) and np.array_equal(out[1][1], ci) and np.array_equal(out[1][2], v)


def test_error_paths(fast, tmp_path):
    assert read(fast, str(tmp_path / "missing.mtx"))[0] == -1
    p = str(tmp_path / "bad.mtx")
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    assert read(fast, p)[0] == -1                       # dense arrays are not supported, as in the reference
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n3 3 2\n1 1 1.0\n")
    assert read(fast, p)[0] == -1                       # fewer entries than announced
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n3 3 1\n1 1 1.0\n")
    assert read(fast, p, need_symm=1)[0] == -1          # need_symm on a general matrix


def test_drivers_link_the_fast_reader():
    exe = os.path.join(PKG, "bin", "test_rp_spmm.exe")
    if not os.path.exists(exe):
        pytest.skip("drivers not built")
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libcrpingest.so" in out
