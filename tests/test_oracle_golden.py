"""Pin the oracle: every integer output and C of oracle/crp_oracle.c must equal, bit for bit,
what the UNMODIFIED reference sources produced (tests/golden/*.npz, see make_golden.py)."""
import os

import numpy as np
import pytest

import cases
import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SKIP = ("C", "layout", "ldC", "C_nrow", "C_ncol")


def load_case(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


@pytest.mark.parametrize("case", cases.SPMM_CASES, ids=[c[0] for c in cases.SPMM_CASES])
def test_oracle_matches_reference(case):
    name, spec, n, mode, nproc, layout, reidx = case
    g = load_case(name)
    m, k, rp, ci, v = cases.build_matrix(spec)
    # the committed CSR is the one the reference ran on
    assert np.array_equal(rp, g["csr_rowptr"]) and np.array_equal(ci, g["csr_colidx"]) and np.array_equal(v, g["csr_val"])
    sim = O.Simulation(m, k, rp, ci, v, n, mode, nproc, layout, reidx)
    Cs = sim.exec()
    for r in range(nproc):
        plan = sim.plan(r)
        keys = [key[len(f"r{r}/"):] for key in g if key.startswith(f"r{r}/")]
        assert keys
        for key in keys:
            ref = g[f"r{r}/{key}"]
            if key in SKIP:
                continue
            if key == "comm_cost" and (r != 0 or mode == "rp"):
                continue            # only rank 0 of a 2-D run holds the cost
            if key == "rA_cost" and r != 0:
                continue
            mine = np.atleast_1d(np.asarray(plan[key]))
            assert mine.shape == ref.shape, (key, r)
            assert np.array_equal(mine.astype(np.float64), ref.astype(np.float64)), (key, r)
        Cref = g[f"r{r}/C"]
        nrow, ncol = int(g[f"r{r}/C_nrow"][0]), int(g[f"r{r}/C_ncol"][0])
        Cref = Cref.reshape(nrow, ncol) if layout == 0 else Cref.reshape(ncol, nrow).T
        assert np.array_equal(Cs[r], Cref), f"C differs on rank {r}"
    sim.close()


def test_known_answers_from_survey():
    """SURVEY.md App. A.1 / A.2 values obtained from the real reference code."""
    assert list(O.row_partition([0, 1, 2, 3, 4, 5, 6, 7], 3)) == [0, 2, 4, 7]
    assert list(O.row_partition([0, 3, 3, 3, 6, 6, 8], 2)) == [0, 4, 6]
    assert list(O.row_partition([0, 2, 4, 6, 8, 8, 8], 2)) == [0, 2, 5]          # trailing empty row lost (reference quirk)
    assert list(O.row_partition([0, 2, 4, 6, 8, 8, 8], 4)) == [0, 1, 2, 3, 5]
    m, k, rp, ci, v = cases.build_matrix(("tridiag", 16))
    rb = np.array([0, 2, 4, 6, 7, 9, 11, 12, 16], np.int32)
    for n, grid, cost in ((1, (8, 1), 14), (4, (8, 1), 56), (16, (4, 2), 165), (64, (2, 4), 335)):
        r = O.part2d(8, m, n, k, rb, rp, ci)
        assert (r["pm"], r["pn"]) == grid and r["comm_cost"] == cost
    r = O.part2d(8, m, 16, k, rb, rp, ci)
    assert list(r["A0_rowptr"]) == [0, 2, 4, 6, 7, 9, 11, 14, 16] and list(r["AC_rowptr"]) == [0, 4, 7, 11, 16] and list(r["BC_colptr"]) == [0, 8, 16]
    r = O.part2d(8, m, 64, k, rb, rp, ci)
    assert list(r["A0_rowptr"]) == [0, 2, 4, 6, 7, 9, 11, 13, 16] and list(r["AC_rowptr"]) == [0, 7, 16] and list(r["BC_colptr"]) == [0, 16, 32, 48, 64]


@pytest.mark.parametrize("name", cases.REDIST_CASES)
def test_oracle_redist_matches_reference(name):
    import ctypes as C
    g = load_case("redist_" + name)
    lay = np.ascontiguousarray(g["layout"], np.int32)
    P = lay.shape[0]
    L = O.lib()
    # plans
    for r in range(P):
        for side, pre in ((0, "s"), (1, "r")):
            ranks, sizes, displs, blks = np.zeros(P, np.int32), np.zeros(P, np.int32), np.zeros(P + 1, np.int32), np.zeros(4 * P, np.int32)
            tot = C.c_int()
            nn = L.orc_redist_plan(P, O.p(lay), r, side, O.p(ranks), O.p(sizes), O.p(displs), O.p(blks), C.byref(tot))
            word = "send" if side == 0 else "recv"
            assert nn == int(g[f"r{r}/n_proc_{word}"][0]) and tot.value == int(g[f"r{r}/{word}_cnt"][0])
            assert np.array_equal(ranks[:nn], g[f"r{r}/{word}_ranks"]) and np.array_equal(sizes[:nn], g[f"r{r}/{word}_sizes"])
            assert np.array_equal(displs[:nn + 1], g[f"r{r}/{word}_displs"]) and np.array_equal(blks[:4 * nn], g[f"r{r}/{pre}blk_sizes"])
    # data
    srcs, dsts, sld, dld = [], [], [], []
    for r in range(P):
        s = lay[r]
        ld_s, ld_d = int(s[3]) + 3, int(s[7]) + 2
        a = np.full((max(int(s[2]), 0), ld_s), -7.0)
        ii, jj = np.meshgrid(np.arange(s[2]), np.arange(s[3]), indexing="ij")
        a[:, :s[3]] = (s[0] + ii) * 1000.5 + (s[1] + jj)
        srcs.append(np.ascontiguousarray(a)); dsts.append(np.full((max(int(s[6]), 0), ld_d), -1.0)); sld.append(ld_s); dld.append(ld_d)
    Sp = (C.c_void_p * P)(*[O.p(a) for a in srcs])
    Dp = (C.c_void_p * P)(*[O.p(a) for a in dsts])
    L.orc_redist_exec(P, O.p(lay), 8, Sp, O.p(O.i32(sld)), Dp, O.p(O.i32(dld)))
    for r in range(P):
        assert np.array_equal(dsts[r].ravel(), g[f"r{r}/dst"]), r
