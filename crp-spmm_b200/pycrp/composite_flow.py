"""The composite engine (include/crpspmm.h) driven like deprecated/examples/test_crpspmm.c:41-141 drives the reference's:
A in a 1-D row layout of the caller's choosing, B and C in an even 2-D block layout over a balanced process grid.

    minimpirun -np P python -m pycrp.composite_flow <csr.bin> <n> <prefix> [--no-exec] [--rowsplit even|nnz] [--gather-c]

Dumps per rank: the grid, the owned-rows CSR after the A redistribution (values included) and, unless --no-exec, the C block.
"""
import argparse
import ctypes as C
import sys

import numpy as np

from . import capi, gen


def even_split(length, nblk):
    base, rem = divmod(length, nblk)
    out = np.zeros(nblk + 1, np.int64)
    for i in range(nblk):
        out[i + 1] = out[i] + base + (1 if i < rem else 0)
    return out


def balanced_dims(nproc):
    best = (nproc, 1)
    for a in range(1, int(nproc ** 0.5) + 1):
        if nproc % a == 0:
            best = (nproc // a, a)
    return best          # non-increasing, like MPI_Dims_create


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("csr")
    ap.add_argument("n", type=int)
    ap.add_argument("prefix")
    ap.add_argument("--no-exec", action="store_true")
    ap.add_argument("--rowsplit", default="even", choices=["even", "skew"])
    ap.add_argument("--gather-c", action="store_true", help="rank 0 wants all of C (the reference driver's check mode)")
    a = ap.parse_args(argv)
    rank, nproc = capi.mpi_init()
    L = capi.load()
    m, k, rowptr, colidx, val = gen.read_csr_bin(a.csr, mmap=True)
    n = a.n
    # A: contiguous row ranges, even or deliberately lopsided
    if a.rowsplit == "even":
        rs = even_split(m, nproc)
    else:
        w = np.arange(1, nproc + 1, dtype=np.float64) ** 2
        rs = np.concatenate([[0], np.round(np.cumsum(w) / w.sum() * m)]).astype(np.int64)
        rs[-1] = m
    a0, a1 = int(rs[rank]), int(rs[rank + 1])
    loc_rp = np.ascontiguousarray(rowptr[a0:a1 + 1], dtype=np.int32)
    loc_ci = np.ascontiguousarray(colidx[rowptr[a0]:rowptr[a1]], dtype=np.int32)
    loc_v = np.ascontiguousarray(val[rowptr[a0]:rowptr[a1]], dtype=np.float64)
    # B, C: even blocks over a balanced grid
    pr, pc = balanced_dims(nproc)
    ri, rj = rank // pc, rank % pc
    bk, bn, cm = even_split(k, pr), even_split(n, pc), even_split(m, pr)
    B_rect = (int(bk[ri]), int(bk[ri + 1] - bk[ri]), int(bn[rj]), int(bn[rj + 1] - bn[rj]))          # srow, nrow, scol, ncol
    C_rect = (int(cm[ri]), int(cm[ri + 1] - cm[ri]), int(bn[rj]), int(bn[rj + 1] - bn[rj]))
    if a.gather_c:
        C_rect = (0, m, 0, n) if rank == 0 else (0, 0, 0, 0)
    eng = C.POINTER(capi.CrpspmmEngine)()
    L.crpspmm_engine_init(m, n, k, a0, a1 - a0, capi.ptr(loc_rp), capi.ptr(loc_ci), B_rect[0], B_rect[1], B_rect[2], B_rect[3],
                          C_rect[0], C_rect[1], C_rect[2], C_rect[3], capi.MPI_COMM_WORLD, 1, C.byref(eng), None)
    e = eng.contents
    out = dict(np_row=e.np_row, np_col=e.np_col, loc_A_srow=e.loc_A_srow, loc_A_nrow=e.loc_A_nrow, loc_A_nnz=e.loc_A_nnz,
               loc_B=np.array([e.loc_B_srow, e.loc_B_nrow, e.loc_B_scol, e.loc_B_ncol]), loc_C=np.array([e.loc_C_srow, e.loc_C_nrow]),
               loc_A_rowptr=capi.np_from(e.loc_A_rowptr, e.loc_A_nrow + 1, np.int32), loc_A_colidx=capi.np_from(e.loc_A_colidx, e.loc_A_nnz, np.int32),
               C_rect=np.array(C_rect))
    if a.no_exec:
        L.crpspmm_engine_redist_A_values(eng, capi.ptr(loc_v))
        out["loc_A_val"] = capi.np_from(e.loc_A_val, e.loc_A_nnz, np.float64)
    else:
        B = np.ascontiguousarray(gen.fill_B(B_rect[0], B_rect[1], B_rect[2], B_rect[3]))
        Cb = np.zeros((C_rect[1], C_rect[3]))
        for _ in range(2):          # second call reuses the engine (values unchanged)
            L.crpspmm_engine_exec(eng, capi.ptr(loc_rp), capi.ptr(loc_ci), capi.ptr(loc_v), capi.ptr(B), max(B_rect[3], 1), capi.ptr(Cb), max(C_rect[3], 1))
        out["C"] = Cb
        out["loc_A_val"] = capi.np_from(e.loc_A_val, e.loc_A_nnz, np.float64)
        out["nelem"] = np.array([e.nelem_A_rd, e.nelem_A_agv, e.nelem_B_rd, e.nelem_B_a2av, e.nelem_B_a2av_min], dtype=np.uint64)
        L.crpspmm_engine_print_stat(eng)
    np.savez(f"{a.prefix}.r{rank}.npz", **out)
    L.crpspmm_engine_free(C.byref(eng))
    capi.mpi_barrier()
    capi.mpi_finalize()
    return 0


if __name__ == "__main__":
    sys.exit(main())
