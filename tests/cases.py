"""Small matrices shared by the golden-vector generator and the tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crp-spmm_b200"))

from pycrp import gen  # noqa: E402


def build_matrix(spec):
    kind = spec[0]
    if kind == "tridiag":
        return gen.tridiag(spec[1])
    if kind == "rand":          # ("rand", m, k, nnz_per_row, seed, empty_rows)
        return gen.random_rect(spec[1], spec[2], spec[3], seed=spec[4], empty_rows=tuple(spec[5]))
    if kind == "stencil27":
        return gen.stencil27(spec[1])
    if kind == "rmat":          # ("rmat", scale, edge_factor, seed)
        m, k, rp, ci, v = gen.rmat(scale=spec[1], edge_factor=spec[2], seed=spec[3])
        return m, k, rp, ci, v
    if kind == "pwtk":          # ("pwtk", m, nnz, bandwidth, grid_w)
        return gen.pwtk_like(m=spec[1], target_nnz=spec[2], bandwidth=spec[3], grid_w=spec[4], seed=7)
    if kind == "blockdiag":     # P dense-ish diagonal blocks with no coupling: zero communication
        nb, bs = spec[1], spec[2]
        rows = np.repeat(np.arange(nb * bs), bs)
        cols = (rows // bs) * bs + np.tile(np.arange(bs), nb * bs)
        vals = 1.0 + 0.001 * (rows * 7 + cols * 3) % 5
        rp, ci, v = gen.coo_to_csr(nb * bs, rows.astype(np.int64), cols.astype(np.int64), vals.astype(np.float64), sum_duplicates=False)
        return nb * bs, nb * bs, rp, ci, v
    raise ValueError(kind)


# name, matrix spec, n, mode, nproc, layout, reidx
SPMM_CASES = [
    ("tridiag8_rp_np2_n4",      ("tridiag", 8),  4,  "rp", 2, 0, 1),
    ("tridiag16_2d_np8_n1",     ("tridiag", 16), 1,  "2d", 8, 0, 1),     # 8x1, cost 14
    ("tridiag16_2d_np8_n4",     ("tridiag", 16), 4,  "2d", 8, 0, 1),     # 8x1, cost 56
    ("tridiag16_2d_np8_n16",    ("tridiag", 16), 16, "2d", 8, 0, 1),     # 4x2, cost 165
    ("tridiag16_2d_np8_n64",    ("tridiag", 16), 64, "2d", 8, 0, 1),     # 2x4, cost 335
    ("rand300_2d_np1_n16",      ("rand", 300, 300, 7, 3, (5, 6, 100)), 16, "2d", 1, 0, 1),
    ("rand300_2d_np2_n16",      ("rand", 300, 300, 7, 3, (5, 6, 100)), 16, "2d", 2, 0, 1),
    ("rand300_2d_np3_n16",      ("rand", 300, 300, 7, 3, (5, 6, 100)), 16, "2d", 3, 0, 1),
    ("rand300_2d_np4_n16",      ("rand", 300, 300, 7, 3, (5, 6, 100)), 16, "2d", 4, 0, 1),
    ("rand300_2d_np6_n24",      ("rand", 300, 300, 7, 3, (5, 6, 100)), 24, "2d", 6, 0, 1),
    ("rand300_2d_np8_n64",      ("rand", 300, 300, 7, 3, (5, 6, 100)), 64, "2d", 8, 0, 1),
    ("rand300_2d_np4_n16_cm",   ("rand", 300, 300, 7, 3, (5, 6, 100)), 16, "2d", 4, 1, 1),     # column-major B / C
    ("rand300_rp_np4_n8",       ("rand", 300, 300, 7, 3, (5, 6, 100)), 8,  "rp", 4, 0, 1),
    ("rand300_rp_np4_n8_noreidx", ("rand", 300, 300, 7, 3, (5, 6, 100)), 8, "rp", 4, 0, 0),   # RP_SPMM_REIDX=0
    ("rand300_rp_np3_n5_cm",    ("rand", 300, 300, 7, 3, (5, 6, 100)), 5,  "rp", 3, 1, 1),     # odd n, column-major
    ("rect200x350_2d_np4_n12",  ("rand", 200, 350, 5, 11, ()), 12, "2d", 4, 0, 1),             # m != k: even B split
    ("rect200x350_rp_np4_n12",  ("rand", 200, 350, 5, 11, ()), 12, "rp", 4, 0, 1),
    ("rect350x200_2d_np6_n32",  ("rand", 350, 200, 9, 12, (0, 349)), 32, "2d", 6, 0, 1),
    ("stencil6_2d_np8_n32",     ("stencil27", 6), 32, "2d", 8, 0, 1),
    ("stencil6_rp_np8_n3",      ("stencil27", 6), 3,  "rp", 8, 0, 1),
    ("rmat8_2d_np8_n16",        ("rmat", 8, 8, 2), 16, "2d", 8, 0, 1),                          # skewed rows
    ("rmat8_rp_np4_n16",        ("rmat", 8, 8, 2), 16, "rp", 4, 0, 1),
    ("pwtk600_2d_np8_n64",      ("pwtk", 600, 30000, 500, 8), 64, "2d", 8, 0, 1),               # pwtk-shaped
    ("pwtk600_2d_np4_n32",      ("pwtk", 600, 30000, 500, 8), 32, "2d", 4, 0, 1),
    ("blockdiag_rp_np4_n8",     ("blockdiag", 4, 8), 8, "rp", 4, 0, 1),                         # zero communication
]

# name, P, global rows, global cols, per-rank (src_srow, src_scol, src_nrow, src_ncol, req_srow, req_scol, req_nrow, req_ncol)
def _grid(P, pr, pc, R, Cc):
    out = []
    for r in range(P):
        i, j = r // pc, r % pc
        rs, re = R * i // pr, R * (i + 1) // pr
        cs, ce = Cc * j // pc, Cc * (j + 1) // pc
        out.append((rs, cs, re - rs, ce - cs))
    return out


def redist_layout(name):
    if name == "rowblk_to_2x2":     # 1-D row blocks -> 2 x 2 grid
        src, req = _grid(4, 4, 1, 37, 22), _grid(4, 2, 2, 37, 22)
    elif name == "2x3_to_3x2":
        src, req = _grid(6, 2, 3, 50, 41), _grid(6, 3, 2, 50, 41)
    elif name == "gather_to_0":     # everything to rank 0 (the drivers' C gather, test_para2d_spmm.c:193-200)
        src = _grid(4, 2, 2, 33, 18)
        req = [(0, 0, 33, 18)] + [(0, 0, 0, 0)] * 3
    elif name == "colblk_to_rowblk_8":
        src, req = _grid(8, 1, 8, 64, 40), _grid(8, 8, 1, 64, 40)
    elif name == "identity_3":
        src = req = _grid(3, 3, 1, 10, 7)
    elif name == "single":
        src = req = [(0, 0, 9, 5)]
    else:
        raise ValueError(name)
    return [s + q for s, q in zip(src, req)]


REDIST_CASES = ["rowblk_to_2x2", "2x3_to_3x2", "gather_to_0", "colblk_to_rowblk_8", "identity_3", "single"]
REDIST_DIMS = {"rowblk_to_2x2": (37, 22), "2x3_to_3x2": (50, 41), "gather_to_0": (33, 18), "colblk_to_rowblk_8": (64, 40),
               "identity_3": (10, 7), "single": (9, 5)}


# Deprecated composite engine (tests/golden/make_golden_crpspmm.py -> tests/golden/crpspmm_tables.json): name, matrix spec, n, nproc
CRPSPMM_CASES = [
    ("cmp_pwtk1500_np1_n32",  ("pwtk", 1500, 77000, 1200, 10), 32,  1),
    ("cmp_pwtk1500_np4_n32",  ("pwtk", 1500, 77000, 1200, 10), 32,  4),     # 4 x 1
    ("cmp_pwtk1500_np6_n512", ("pwtk", 1500, 77000, 1200, 10), 512, 6),     # 1 x 6
    ("cmp_pwtk1500_np8_n128", ("pwtk", 1500, 77000, 1200, 10), 128, 8),
    ("cmp_rand300_np4_n64",   ("rand", 300, 300, 7, 3, ()), 64, 4),
]
