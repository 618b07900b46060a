"""Synthetic matrix generators for the BASELINE.json configs and
the binary CSR / Matrix Market writers shared by the library drivers, the
oracle tools and bench.py.  The reference itself has no generators (it reads
SuiteSparse .mtx files, examples/test_utils.c:21-55) and no RNG; B is always
B[i, j] = 0.19 i + 0.24 j on global indices (examples/test_utils.c:121-154).

All generators return (m, k, rowptr[int32], colidx[int32], val[float64]) with
sorted, duplicate-free column indices per row.
"""
from __future__ import annotations

import numpy as np

MAGIC = b"CRPCSR1\0"


# --------------------------------------------------------------------- helpers
def coo_to_csr(m, rows, cols, vals, sum_duplicates=True):
    """Sort COO by (row, col); optionally merge duplicates by summation."""
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    if sum_duplicates and rows.size:
        key_change = np.empty(rows.size, dtype=bool)
        key_change[0] = True
        key_change[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
        if not key_change.all():
            grp = np.cumsum(key_change) - 1
            vals = np.bincount(grp, weights=vals, minlength=int(grp[-1]) + 1)
            rows, cols = rows[key_change], cols[key_change]
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.add.at(rowptr, rows.astype(np.int64) + 1, 1) if rows.size < 1 << 16 else None
    if rows.size >= 1 << 16:
        rowptr[1:] = np.bincount(rows, minlength=m)
    rowptr = np.cumsum(rowptr)
    assert rowptr[-1] < 2 ** 31
    return rowptr.astype(np.int32), cols.astype(np.int32), vals.astype(np.float64)


def fill_B(srow, nrow, scol, ncol, dtype=np.float64):
    """B[i, j] = 0.19 * i + 0.24 * j on global indices, evaluated in fp64 exactly
    as examples/test_utils.c:133-141 does (int * double + int * double)."""
    i = np.arange(srow, srow + nrow, dtype=np.float64)[:, None] * 0.19
    j = np.arange(scol, scol + ncol, dtype=np.float64)[None, :] * 0.24
    return (i + j).astype(dtype)


# ------------------------------------------------------------------ generators
def pwtk_like(m=217918, target_nnz=11634424, bandwidth=189331, grid_w=64, dof=6, seed=20231117):
    """pwtk-shaped SPD matrix: a shell-type FEM stiffness pattern.

    SuiteSparse pwtk (217,918 rows, 11,634,424 nnz, ~53.4 nnz/row, bandwidth
    189,331) is a pressurised wind-tunnel shell model: 6 unknowns per node,
    nodes coupled to their mesh neighbours.  We reproduce that shape with a
    structured quad mesh `grid_w` nodes wide: every node couples to its 8 mesh
    neighbours and itself with dense dof x dof blocks (54 nnz per interior row,
    narrow band), plus a few long-range node couplings (stiffeners) so that
    max|i-j| equals `bandwidth` and nnz lands within 0.1 % of `target_nnz`.
    Off-diagonals are uniform(-1, 0), the diagonal is 1 + sum|offdiag| (strictly
    diagonally dominant, symmetric => SPD).  PRNG: PCG64 via default_rng(seed).
    """
    rng = np.random.default_rng(seed)
    nn = -(-m // dof)                                   # nodes; the last one may be short
    node_lo = np.arange(nn, dtype=np.int64) * dof
    node_sz = np.minimum(node_lo + dof, m) - node_lo
    gx = np.arange(nn) % grid_w
    # forward mesh neighbours of node u: +1 (same row), +W-1, +W, +W+1 (next row)
    pairs_u, pairs_v = [], []
    u = np.arange(nn, dtype=np.int64)
    for off, ok in ((1, gx < grid_w - 1), (grid_w - 1, gx > 0), (grid_w, np.ones(nn, bool)), (grid_w + 1, gx < grid_w - 1)):
        v = u + off
        keep = ok & (v < nn)
        pairs_u.append(u[keep]); pairs_v.append(v[keep])
    pu = np.concatenate(pairs_u); pv = np.concatenate(pairs_v)

    def block_nnz(a, b):
        return int((node_sz[a] * node_sz[b]).sum())

    nnz_now = int((node_sz * node_sz).sum()) + 2 * block_nnz(pu, pv)
    # long-range couplings: one pinned so that the bandwidth is exact, the rest random
    far_u, far_v = [], []
    if bandwidth is not None and bandwidth < m and bandwidth > (grid_w + 2) * dof:
        i0 = 0
        j0 = bandwidth                                     # entry (0-th dof of node 0, dof j0)
        far_u.append(0); far_v.append(j0 // dof)
    extra_pairs = max(0, (target_nnz - nnz_now) // (2 * dof * dof)) if target_nnz else 0
    max_node_dist = (bandwidth // dof - 1) if bandwidth else nn // 2
    if extra_pairs > len(far_u):
        cnt = int(extra_pairs - len(far_u))
        a = rng.integers(0, nn - 1, size=cnt)
        d = rng.integers(grid_w + 2, max(grid_w + 3, max_node_dist), size=cnt)
        b = a + d
        keep = b < nn - 1
        far_u.extend(a[keep].tolist()); far_v.extend(b[keep].tolist())
    if far_u:
        fu = np.asarray(far_u, dtype=np.int64); fv = np.asarray(far_v, dtype=np.int64)
        # drop accidental duplicates of existing pairs
        key_far = fu * nn + fv
        _, first = np.unique(key_far, return_index=True)
        fu, fv = fu[np.sort(first)], fv[np.sort(first)]
        pu = np.concatenate([pu, fu]); pv = np.concatenate([pv, fv])

    # expand node pairs (u < v) into dense dof x dof blocks (upper part), then mirror
    szu, szv = node_sz[pu], node_sz[pv]
    blk = szu * szv
    tot = int(blk.sum())
    pair_id = np.repeat(np.arange(pu.size), blk)
    inner = np.arange(tot) - np.repeat(np.cumsum(blk) - blk, blk)
    r_up = node_lo[pu][pair_id] + inner // szv[pair_id]
    c_up = node_lo[pv][pair_id] + inner % szv[pair_id]
    # strict upper triangle of the diagonal blocks
    dblk = node_sz * (node_sz - 1) // 2
    iu, ju = np.triu_indices(dof, 1)
    dr = (node_lo[:, None] + iu[None, :]).ravel()
    dc = (node_lo[:, None] + ju[None, :]).ravel()
    keep = (dr < m) & (dc < m)
    r_up = np.concatenate([r_up, dr[keep]]); c_up = np.concatenate([c_up, dc[keep]])
    if bandwidth is not None and far_u:
        # the pinned far block of node pair (0, bandwidth//dof) spans columns up to
        # node_lo + dof - 1; trim it so that max|i-j| is exactly `bandwidth`
        too_far = (c_up - r_up) > bandwidth
        r_up, c_up = r_up[~too_far], c_up[~too_far]
    v_up = -rng.random(r_up.size)
    rows = np.concatenate([r_up, c_up, np.arange(m)])
    cols = np.concatenate([c_up, r_up, np.arange(m)])
    absrow = np.bincount(r_up, weights=-v_up, minlength=m) + np.bincount(c_up, weights=-v_up, minlength=m)
    vals = np.concatenate([v_up, v_up, 1.0 + absrow])
    rowptr, colidx, val = coo_to_csr(m, rows, cols, vals, sum_duplicates=False)
    return m, m, rowptr, colidx, val


def erdos_renyi(scale=22, nnz_per_row=16, seed=1):
    """Uniform random: exactly `nnz_per_row` distinct uniform columns per row, values uniform(-1, 1)."""
    rng = np.random.default_rng(seed)
    m = 1 << scale
    cols = rng.integers(0, m, size=(m, nnz_per_row), dtype=np.int64)
    cols.sort(axis=1)
    while True:
        dup_rows = np.nonzero((cols[:, 1:] == cols[:, :-1]).any(axis=1))[0]
        if dup_rows.size == 0:
            break
        cols[dup_rows] = np.sort(rng.integers(0, m, size=(dup_rows.size, nnz_per_row), dtype=np.int64), axis=1)
    val = rng.uniform(-1.0, 1.0, size=m * nnz_per_row)
    rowptr = (np.arange(m + 1, dtype=np.int64) * nnz_per_row).astype(np.int32)
    return m, m, rowptr, cols.ravel().astype(np.int32), val


def rmat(scale=22, edge_factor=32, abcd=(0.57, 0.19, 0.19, 0.05), seed=2, chunk=1 << 24):
    """Graph500 R-MAT, duplicates merged (summed), no vertex permutation, values uniform(0, 1),
    plus a unit diagonal so that no row is empty."""
    rng = np.random.default_rng(seed)
    m = 1 << scale
    ne = m * edge_factor
    a, b, c, _ = abcd
    rows_l, cols_l, vals_l = [], [], []
    for s in range(0, ne, chunk):
        cnt = min(chunk, ne - s)
        r = np.zeros(cnt, dtype=np.int64); cidx = np.zeros(cnt, dtype=np.int64)
        for _bit in range(scale):
            u = rng.random(cnt)
            rbit = u >= a + b
            cbit = ((u >= a) & (u < a + b)) | (u >= a + b + c)
            r = (r << 1) | rbit
            cidx = (cidx << 1) | cbit
        rows_l.append(r); cols_l.append(cidx); vals_l.append(rng.random(cnt))
    rows = np.concatenate(rows_l + [np.arange(m, dtype=np.int64)])
    cols = np.concatenate(cols_l + [np.arange(m, dtype=np.int64)])
    vals = np.concatenate(vals_l + [np.ones(m)])
    rowptr, colidx, val = coo_to_csr(m, rows, cols, vals, sum_duplicates=True)
    return m, m, rowptr, colidx, val


def stencil27(n=128, centre=27.0):
    """3-D 27-point stencil on an n^3 grid, non-periodic, centre `centre`, neighbours -1."""
    idx = np.arange(n ** 3, dtype=np.int64).reshape(n, n, n)
    rows_l, cols_l, vals_l = [], [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                zs = slice(max(0, -dz), n - max(0, dz)); zd = slice(max(0, dz), n - max(0, -dz))
                ys = slice(max(0, -dy), n - max(0, dy)); yd = slice(max(0, dy), n - max(0, -dy))
                xs = slice(max(0, -dx), n - max(0, dx)); xd = slice(max(0, dx), n - max(0, -dx))
                r = idx[zs, ys, xs].ravel(); c = idx[zd, yd, xd].ravel()
                rows_l.append(r); cols_l.append(c)
                vals_l.append(np.full(r.size, centre if (dx == 0 and dy == 0 and dz == 0) else -1.0))
    m = n ** 3
    rowptr, colidx, val = coo_to_csr(m, np.concatenate(rows_l), np.concatenate(cols_l), np.concatenate(vals_l), sum_duplicates=False)
    return m, m, rowptr, colidx, val


def random_rect(m, k, nnz_per_row, seed=0, empty_rows=()):
    """Small general test matrix (m != k allowed, some rows may be empty)."""
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    for i in range(m):
        if i in empty_rows:
            continue
        cnt = int(rng.integers(1, max(2, min(k, 2 * nnz_per_row))))
        cs = rng.choice(k, size=min(cnt, k), replace=False)
        rows.append(np.full(cs.size, i)); cols.append(cs)
    rows = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    vals = rng.uniform(-1.0, 1.0, size=rows.size)
    rowptr, colidx, val = coo_to_csr(m, rows.astype(np.int64), cols.astype(np.int64), vals, sum_duplicates=False)
    return m, k, rowptr, colidx, val


def tridiag(m):
    rows = np.concatenate([np.arange(m), np.arange(1, m), np.arange(m - 1)])
    cols = np.concatenate([np.arange(m), np.arange(m - 1), np.arange(1, m)])
    vals = np.concatenate([np.full(m, 2.0), np.full(m - 1, -1.0), np.full(m - 1, -1.0)])
    rowptr, colidx, val = coo_to_csr(m, rows.astype(np.int64), cols.astype(np.int64), vals, sum_duplicates=False)
    return m, m, rowptr, colidx, val


# ------------------------------------------------------------------------- I/O
def write_csr_bin(path, m, k, rowptr, colidx, val):
    with open(path, "wb") as f:
        f.write(MAGIC)
        np.asarray([m, k, int(rowptr[-1])], dtype=np.int64).tofile(f)
        np.ascontiguousarray(rowptr, dtype=np.int32).tofile(f)
        np.ascontiguousarray(colidx, dtype=np.int32).tofile(f)
        np.ascontiguousarray(val, dtype=np.float64).tofile(f)


def read_csr_bin(path, mmap=False):
    """(m, k, rowptr, colidx, val); with mmap=True the three arrays are read-only maps of the file."""
    with open(path, "rb") as f:
        assert f.read(8) == MAGIC
        m, k, nnz = (int(x) for x in np.fromfile(f, dtype=np.int64, count=3))
        if mmap:
            off = 32
            rowptr = np.memmap(path, dtype=np.int32, mode="r", offset=off, shape=(m + 1,))
            off += 4 * (m + 1)
            colidx = np.memmap(path, dtype=np.int32, mode="r", offset=off, shape=(nnz,)) if nnz else np.zeros(0, np.int32)
            off += 4 * nnz
            val = np.memmap(path, dtype=np.float64, mode="r", offset=off, shape=(nnz,)) if nnz else np.zeros(0, np.float64)
            return m, k, rowptr, colidx, val
        rowptr = np.fromfile(f, dtype=np.int32, count=m + 1)
        colidx = np.fromfile(f, dtype=np.int32, count=nnz)
        val = np.fromfile(f, dtype=np.float64, count=nnz)
    return m, k, rowptr, colidx, val


def write_mtx(path, m, k, rowptr, colidx, val):
    """Matrix Market 'coordinate real general', 1-based, full precision (%.17g)."""
    rows = np.repeat(np.arange(m, dtype=np.int64), np.diff(rowptr)) + 1
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{m} {k} {int(rowptr[-1])}\n")
        step = 1 << 20
        for s in range(0, rows.size, step):
            e = min(rows.size, s + step)
            lines = [f"{r} {c} {v:.17g}\n" for r, c, v in zip(rows[s:e].tolist(), (colidx[s:e].astype(np.int64) + 1).tolist(), val[s:e].tolist())]
            f.write("".join(lines))


def read_dump(path):
    """Read the record stream written by oracle/ref_dump.c / the library's dump helpers."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        name = data[pos:pos + 24].split(b"\0")[0].decode(); pos += 24
        esz, cnt = np.frombuffer(data, dtype=np.int64, count=2, offset=pos); pos += 16
        esz, cnt = int(esz), int(cnt)
        if name in ("C", "A_val", "dst"):
            dt = np.float64
        elif esz == 8:
            dt = np.uint64
        else:
            dt = np.int32
        out[name] = np.frombuffer(data, dtype=dt, count=cnt, offset=pos).copy()
        pos += esz * cnt
    return out
