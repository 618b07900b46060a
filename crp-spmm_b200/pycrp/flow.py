"""The reference drivers' flow, step by step, through the C-ABI of libcrpspmm.so.

    mode "2d": examples/test_para2d_spmm.c:41-167 (1-D nnz split on rank 0 ->
               calc_spmm_part2d_from_1d -> broadcast -> rows scattered by A0_rowptr with
               GLOBAL nnz offsets in rowptr -> fill_B -> para2d_spmm_init -> exec)
    mode "rp": examples/test_rp_spmm.c:41-145 (1-D split, x_displs = row split if square)

Used by the tests (under `minimpirun -np P python -m pycrp.flow ...`), by bench.py
and by __graft_entry__.smoke().  It reads a binary CSR (pycrp.gen.write_csr_bin)
instead of a .mtx file; every rank maps the file and takes its slice.
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

import numpy as np

from . import capi, gen


class Problem:
    """One rank's view of a distributed SpMM set up like the reference drivers do."""

    def __init__(self, csr_path, n, mode="2d", layout=0, dtype=np.float64, rank=None, nproc=None, comm=capi.MPI_COMM_WORLD):
        L = capi.load()
        self.L, self.n, self.mode, self.layout, self.dtype, self.comm = L, int(n), mode, int(layout), np.dtype(dtype), comm
        if rank is None:
            r, s = C.c_int(), C.c_int()
            capi.mpi().MPI_Comm_rank(comm, C.byref(r))
            capi.mpi().MPI_Comm_size(comm, C.byref(s))
            rank, nproc = r.value, s.value
        self.rank, self.nproc = rank, nproc
        m, k, rowptr, colidx, val = gen.read_csr_bin(csr_path, mmap=True)
        self.m, self.k, self.nnz = m, k, int(rowptr[m])
        self.rowptr_g = rowptr

        # ---- rank 0 plans, everybody learns the result ----
        hdr = np.zeros(4, dtype=np.int64)                 # pm, pn, comm_cost, (unused)
        rb = np.zeros(nproc + 1, dtype=np.int32)
        self.t_part = 0.0
        if rank == 0:
            rp32 = np.ascontiguousarray(rowptr, dtype=np.int32)
            t0 = L.get_wtime_sec()
            L.csr_mat_row_partition(m, capi.ptr(rp32), nproc, capi.ptr(rb))
            hdr[0], hdr[1] = nproc, 1
            if mode == "2d":
                ci32 = np.ascontiguousarray(colidx, dtype=np.int32)
                pm, pn, cost = C.c_int(), C.c_int(), C.c_size_t()
                a0, br, ac, bc = capi.c_int_p(), capi.c_int_p(), capi.c_int_p(), capi.c_int_p()
                L.calc_spmm_part2d_from_1d(nproc, m, self.n, k, capi.ptr(rb), capi.ptr(rp32), capi.ptr(ci32), 1,
                                           C.byref(pm), C.byref(pn), C.byref(cost), C.byref(a0), C.byref(br), C.byref(ac), C.byref(bc), 0)
                hdr[0], hdr[1], hdr[2] = pm.value, pn.value, cost.value
                splits = [capi.np_from(a0, nproc + 1, np.int32), capi.np_from(br, pm.value + 1, np.int32),
                          capi.np_from(ac, pm.value + 1, np.int32), capi.np_from(bc, pn.value + 1, np.int32)]
                libc = C.CDLL(None)
                libc.free.argtypes = [C.c_void_p]
                for p_ in (a0, br, ac, bc):
                    libc.free(C.cast(p_, C.c_void_p))
                del ci32
            self.t_part = L.get_wtime_sec() - t0
        capi.mpi_bcast(hdr, 0, comm)
        capi.mpi_bcast(rb, 0, comm)
        self.pm, self.pn, self.comm_cost = int(hdr[0]), int(hdr[1]), int(hdr[2])
        self.rb_displs0 = rb
        pm, pn = self.pm, self.pn
        if mode == "2d":
            if rank != 0:
                splits = [np.zeros(nproc + 1, np.int32), np.zeros(pm + 1, np.int32), np.zeros(pm + 1, np.int32), np.zeros(pn + 1, np.int32)]
            for a in splits:
                capi.mpi_bcast(a, 0, comm)
            self.A0_rowptr, self.B_rowptr, self.AC_rowptr, self.BC_colptr = splits
            split = self.A0_rowptr
        else:
            split = rb

        # ---- my rows of A; the row-pointer slice keeps GLOBAL nnz offsets (examples/test_utils.c:78-91) ----
        self.a_srow, self.a_nrow = int(split[rank]), int(split[rank + 1] - split[rank])
        z0, z1 = int(rowptr[self.a_srow]), int(rowptr[self.a_srow + self.a_nrow])
        self.loc_rowptr = np.ascontiguousarray(rowptr[self.a_srow:self.a_srow + self.a_nrow + 1], dtype=np.int32)
        self.loc_colidx = np.ascontiguousarray(colidx[z0:z1], dtype=np.int32)
        self.loc_val = np.ascontiguousarray(val[z0:z1], dtype=np.float64)

        # ---- my B and C blocks ----
        if mode == "2d":
            pi, pj = rank // pn, rank % pn
            self.b_srow, self.b_nrow = int(self.B_rowptr[pi]), int(self.B_rowptr[pi + 1] - self.B_rowptr[pi])
            self.c_srow, self.c_nrow = int(self.AC_rowptr[pi]), int(self.AC_rowptr[pi + 1] - self.AC_rowptr[pi])
            self.bc_scol, self.bc_ncol = int(self.BC_colptr[pj]), int(self.BC_colptr[pj + 1] - self.BC_colptr[pj])
            self.x_displs = None
        else:
            if m == k:
                self.x_displs = rb.copy()
            else:
                self.x_displs = np.zeros(nproc + 1, np.int32)
                sp, sz = C.c_int(), C.c_int()
                for i in range(nproc + 1):
                    L.calc_block_spos_size(k, nproc, i, C.byref(sp), C.byref(sz))
                    self.x_displs[i] = sp.value
            self.b_srow, self.b_nrow = int(self.x_displs[rank]), int(self.x_displs[rank + 1] - self.x_displs[rank])
            self.c_srow, self.c_nrow = self.a_srow, self.a_nrow
            self.bc_scol, self.bc_ncol = 0, self.n
        self.ldB = self.bc_ncol if layout == 0 else self.b_nrow
        self.ldC = self.bc_ncol if layout == 0 else self.c_nrow
        self.p2d = None
        self.rp = None

    # -- B / C ---------------------------------------------------------------
    def make_B(self):
        """B[i, j] = 0.19 i + 0.24 j on global indices (examples/test_utils.c:121-154)."""
        B = gen.fill_B(self.b_srow, self.b_nrow, self.bc_scol, self.bc_ncol, dtype=self.dtype)
        return np.ascontiguousarray(B if self.layout == 0 else B.T)     # column-major == transposed row-major

    def empty_C(self):
        shape = (self.c_nrow, self.bc_ncol) if self.layout == 0 else (self.bc_ncol, self.c_nrow)
        return np.zeros(shape, dtype=self.dtype)

    def C_rowmajor(self, C_):
        return C_ if self.layout == 0 else np.ascontiguousarray(C_.T)

    # -- engine ---------------------------------------------------------------
    def init(self):
        L = self.L
        if self.mode == "2d":
            self.p2d = C.POINTER(capi.Para2dSpmm)()
            L.para2d_spmm_init(self.comm, self.pm, self.pn, capi.ptr(self.A0_rowptr), capi.ptr(self.B_rowptr), capi.ptr(self.AC_rowptr),
                               capi.ptr(self.BC_colptr), capi.ptr(self.loc_rowptr), capi.ptr(self.loc_colidx), capi.ptr(self.loc_val), C.byref(self.p2d))
            self.rp = self.p2d.contents.rp_spmm
        else:
            self.rp = C.POINTER(capi.RowparaSpmm)()
            L.rp_spmm_init(self.a_srow, self.a_nrow, capi.ptr(self.loc_rowptr), capi.ptr(self.loc_colidx), capi.ptr(self.loc_val),
                           capi.ptr(self.x_displs), self.n, self.comm, C.byref(self.rp))
        return self

    def exec_ptr(self, B_ptr, C_ptr):
        """One exec on raw pointers (host or device)."""
        f32 = self.dtype == np.float32
        if self.mode == "2d":
            (self.L.para2d_spmm_exec_f32 if f32 else self.L.para2d_spmm_exec)(self.p2d, self.layout, B_ptr, self.ldB, C_ptr, self.ldC)
        else:
            (self.L.rp_spmm_exec_f32 if f32 else self.L.rp_spmm_exec)(self.rp, self.layout, B_ptr, self.ldB, C_ptr, self.ldC)

    def exec_host(self, B, C_):
        self.exec_ptr(capi.ptr(B), capi.ptr(C_))

    def clear_stat(self):
        self.L.rp_spmm_clear_stat(self.rp)

    def print_stat(self):
        if self.mode == "2d":
            self.L.para2d_spmm_print_stat(self.p2d)
        else:
            self.L.rp_spmm_print_stat(self.rp)

    def plan_dict(self):
        d = capi.rp_plan_dict(self.rp)
        d.update(pm=self.pm, pn=self.pn, comm_cost=self.comm_cost, rb_displs0=self.rb_displs0)
        if self.mode == "2d":
            d.update(A0_rowptr=self.A0_rowptr, B_rowptr=self.B_rowptr, AC_rowptr=self.AC_rowptr, BC_colptr=self.BC_colptr,
                     rA_cost=int(self.p2d.contents.rA_cost))
        else:
            d.update(x_displs=self.x_displs)
        return d

    def algorithmic_bytes(self):
        """4 (A_nrow + 1) + (4 + s) nnz_loc + s rB_nrow n + s A_nrow n for this rank's local product."""
        r = self.rp.contents
        s = self.dtype.itemsize
        nnz_loc = int(r.A_rowptr[r.A_nrow])
        return 4 * (r.A_nrow + 1) + (4 + s) * nnz_loc + s * r.rB_nrow * r.glb_n + s * r.A_nrow * r.glb_n, 2 * nnz_loc * r.glb_n

    def free(self):
        if self.p2d is not None:
            self.L.para2d_spmm_free(C.byref(self.p2d))
            self.p2d = None
        elif self.rp is not None:
            self.L.rp_spmm_free(C.byref(self.rp))
        self.rp = None


def main(argv=None):
    ap = argparse.ArgumentParser(description="driver flow through libcrpspmm.so; dumps the plan (and C) per rank as .npz")
    ap.add_argument("csr")
    ap.add_argument("n", type=int)
    ap.add_argument("mode", choices=["2d", "rp"])
    ap.add_argument("--layout", type=int, default=0)
    ap.add_argument("--dump", default=None, help="prefix of <prefix>.r<rank>.npz")
    ap.add_argument("--no-exec", action="store_true", help="plan only (CRP_SPMM_PLAN_ONLY=1 on a box without GPU)")
    ap.add_argument("--device", action="store_true", help="B and C resident on the device (zero-copy exec)")
    ap.add_argument("--f32", action="store_true")
    ap.add_argument("--ntest", type=int, default=1)
    ap.add_argument("--stat", action="store_true")
    a = ap.parse_args(argv)
    rank, nproc = capi.mpi_init()
    pb = Problem(a.csr, a.n, a.mode, a.layout, np.float32 if a.f32 else np.float64, rank, nproc).init()
    out = pb.plan_dict()
    if not a.no_exec:
        B, C_ = pb.make_B(), pb.empty_C()
        if a.device:
            dB, dC = capi.DevBuf.from_numpy(B), capi.DevBuf(C_.nbytes)
            for _ in range(a.ntest):
                pb.exec_ptr(dB.p, dC.p)
            C_ = dC.to_numpy(C_.shape, C_.dtype)
            dB.free(), dC.free()
        else:
            for _ in range(a.ntest):
                pb.exec_host(B, C_)
        out["C"] = pb.C_rowmajor(C_)
        out["kernel"] = np.array(pb.L.rp_spmm_kernel_name(pb.rp).decode())
        out["transport"] = np.array(pb.L.rp_spmm_transport_name(pb.rp).decode())
        if a.stat:
            pb.print_stat()
    if a.dump:
        np.savez(f"{a.dump}.r{rank}.npz", **out)
    pb.free()
    capi.mpi_barrier()
    capi.mpi_finalize()
    return 0


if __name__ == "__main__":
    sys.exit(main())
