"""The composite engine (include/crpspmm.h) driven like deprecated/examples/test_crpspmm.c:41-141 drives the reference's:
A in a 1-D row layout of the caller's choosing, B and C in an even 2-D block layout over a balanced process grid.

    minimpirun -np P python -m pycrp.composite_flow <csr.bin> <n> <prefix> [--no-exec] [--rowsplit even|nnz] [--gather-c]
                                                   [--device] [--ntest N] [--json out.json] [--no-dump]
--device keeps the caller's B and C blocks on the GPU (the redistributions then run device to device over NCCL / NVLink);
--json makes rank 0 write the phase times of the timed execs and the redistribution bandwidth against the NVLink peak.

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
    ap.add_argument("--device", action="store_true", help="the caller's B and C blocks are device buffers")
    ap.add_argument("--ntest", type=int, default=2, help="number of execs (the first one builds the 2-D engine and is not timed)")
    ap.add_argument("--json", default="", help="rank 0: write phase times / bandwidths of the timed execs here")
    ap.add_argument("--no-dump", action="store_true")
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
        import json
        import time
        B = np.ascontiguousarray(gen.fill_B(B_rect[0], B_rect[1], B_rect[2], B_rect[3]))
        Cb = np.zeros((C_rect[1], C_rect[3]))
        if a.device:
            dB, dC = capi.DevBuf.from_numpy(B), capi.DevBuf(max(Cb.nbytes, 8))
            pB, pC = dB.p, dC.p
        else:
            pB, pC = capi.ptr(B), capi.ptr(Cb)
        # the first call builds the 2-D engine (A redistribution + replicate-A) and is not timed; the later ones reuse it
        L.crpspmm_engine_exec(eng, capi.ptr(loc_rp), capi.ptr(loc_ci), capi.ptr(loc_v), pB, max(B_rect[3], 1), pC, max(C_rect[3], 1))
        t_first = dict(t_rd_A=e.t_rd_A, t_agv_A=e.t_agv_A, t_exec=e.t_exec)
        L.crpspmm_engine_clear_stat(eng)
        capi.mpi_barrier()
        t0 = time.time()
        for _ in range(max(a.ntest - 1, 1)):
            L.crpspmm_engine_exec(eng, capi.ptr(loc_rp), capi.ptr(loc_ci), capi.ptr(loc_v), pB, max(B_rect[3], 1), pC, max(C_rect[3], 1))
        capi.mpi_barrier()
        wall = (time.time() - t0) / max(a.ntest - 1, 1)
        if a.device:
            Cb = dC.to_numpy(Cb.shape, np.float64) if Cb.size else Cb
        out["C"] = Cb
        out["loc_A_val"] = capi.np_from(e.loc_A_val, e.loc_A_nnz, np.float64)
        out["nelem"] = np.array([e.nelem_A_rd, e.nelem_A_agv, e.nelem_B_rd, e.nelem_B_a2av, e.nelem_B_a2av_min], dtype=np.uint64)
        nx = max(e.n_exec, 1)
        # bytes that really cross between GPUs in the two redistributions: what this rank receives from others
        rdB, rdC = e.rd_B.contents, e.rd_C.contents
        def foreign(rd):
            tot = 0
            for i in range(rd.n_proc_recv):
                if rd.recv_ranks[i] != rank:
                    tot += int(rd.recv_sizes[i])
            return 8 * tot
        mine = dict(rank=rank, t_rd_B=e.t_rd_B / nx, t_rd_C=e.t_rd_C / nx, t_a2a_B=e.t_a2a_B / nx, t_spmm=e.t_spmm / nx, t_exec_nr=e.t_exec_nr / nx,
                    t_exec=e.t_exec / nx, rd_B_recv_bytes=foreign(rdB), rd_C_recv_bytes=foreign(rdC), wall=wall, first=t_first)
        L.crpspmm_engine_print_stat(eng)
        if a.json:
            with open(f"{a.json}.r{rank}", "w") as f:
                json.dump(mine, f)
            capi.mpi_barrier()
            if rank == 0:
                rows = [json.load(open(f"{a.json}.r{q}")) for q in range(nproc)]
                peak = 770.0        # GB/s, measured peer copy (B200_PROFILING.md)
                def bw(key_b, key_t):
                    return max((r[key_b] / max(r[key_t], 1e-12) / 1e9) for r in rows)
                summ = dict(nproc=nproc, grid=[int(e.np_row), int(e.np_col)], n=n, m=int(m), nnz=int(rowptr[-1]), device=bool(a.device),
                            ms_per_exec=1e3 * max(r["wall"] for r in rows),
                            phases_ms_max={k_: 1e3 * max(r[k_] for r in rows) for k_ in ("t_rd_B", "t_a2a_B", "t_spmm", "t_exec_nr", "t_rd_C", "t_exec")},
                            first_exec_s={k_: max(r["first"][k_] for r in rows) for k_ in ("t_rd_A", "t_agv_A", "t_exec")},
                            redist_B={"recv_MB_max": max(r["rd_B_recv_bytes"] for r in rows) / 1e6, "GBps_best_rank": bw("rd_B_recv_bytes", "t_rd_B"),
                                      "nvlink_frac": bw("rd_B_recv_bytes", "t_rd_B") / peak},
                            redist_C={"recv_MB_max": max(r["rd_C_recv_bytes"] for r in rows) / 1e6, "GBps_best_rank": bw("rd_C_recv_bytes", "t_rd_C"),
                                      "nvlink_frac": bw("rd_C_recv_bytes", "t_rd_C") / peak},
                            gflops=2.0 * float(rowptr[-1]) * n / max(r["wall"] for r in rows) / 1e9, per_rank=rows)
                with open(a.json, "w") as f:
                    json.dump(summ, f)
                print("COMPOSITE " + json.dumps({k_: v_ for k_, v_ in summ.items() if k_ != "per_rank"}), flush=True)
    if a.no_dump:
        out = {}
    if not a.no_dump:
        np.savez(f"{a.prefix}.r{rank}.npz", **out)
    L.crpspmm_engine_free(C.byref(eng))
    capi.mpi_barrier()
    capi.mpi_finalize()
    return 0


if __name__ == "__main__":
    sys.exit(main())
