"""ctypes access to oracle/liboracle.so (the CPU restatement) for the tests.
Test-side only: the product never loads this library."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_orc = None
ip = C.POINTER(C.c_int)
dp = C.POINTER(C.c_double)


class OrcRp(C.Structure):
    _fields_ = [
        ("nproc", C.c_int), ("my_rank", C.c_int), ("glb_n", C.c_int), ("A_nrow", C.c_int), ("rB_nrow", C.c_int),
        ("rB_self_src_offset", C.c_int), ("rB_self_dst_offset", C.c_int), ("rB_self_nrow", C.c_int), ("rB_reidx", C.c_int),
        ("A_rowptr", ip), ("A_colidx", ip), ("rB_self_src_ridxs", ip),
        ("rB_scnts", ip), ("rB_sridxs", ip), ("rB_sdispls", ip), ("rB_rcnts", ip), ("rB_rridxs", ip), ("rB_rdispls", ip),
        ("A_val", dp), ("rB_recv_size", C.c_uint64), ("rB_srow", C.c_int), ("rowmap", ip),
    ]


def lib():
    global _orc
    if _orc is None:
        _orc = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        _orc.orc_rp_create.restype = C.POINTER(OrcRp)
        _orc.orc_rp_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _orc.orc_rp_link.argtypes = [C.POINTER(C.POINTER(OrcRp)), C.c_int, C.c_void_p]
        _orc.orc_rp_free.argtypes = [C.POINTER(OrcRp)]
        _orc.orc_rp_exec.argtypes = [C.POINTER(C.POINTER(OrcRp)), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _orc.orc_para2d_rA_cost.restype = C.c_uint64
        _orc.orc_redist_plan.restype = C.c_int
    return _orc


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def arr(ptr, n, dt=np.int32):
    if n <= 0:
        return np.zeros(0, dt)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True)


def row_partition(rowptr, nblk):
    rowptr = i32(rowptr)
    out = np.zeros(nblk + 1, np.int32)
    lib().orc_row_partition(len(rowptr) - 1, p(rowptr), nblk, p(out))
    return out


def block_split(length, nblk):
    out = np.zeros(nblk + 1, np.int32)
    sp, sz = C.c_int(), C.c_int()
    for i in range(nblk + 1):
        lib().orc_block_spos_size(length, nblk, i, C.byref(sp), C.byref(sz))
        out[i] = sp.value
    return out


def comm_size(m, k, rowptr, colidx, rblk, xd):
    rowptr, colidx, rblk, xd = i32(rowptr), i32(colidx), i32(rblk), i32(xd)
    nblk = len(rblk) - 1
    sizes = np.zeros(nblk, np.int32)
    tot = C.c_int()
    lib().orc_comm_size(m, k, p(rowptr), p(colidx), nblk, p(rblk), p(xd), p(sizes), C.byref(tot))
    return sizes, tot.value


def part2d(nproc, m, n, k, rb, rowptr, colidx, rA=1):
    rb, rowptr, colidx = i32(rb), i32(rowptr), i32(colidx)
    pm, pn, cost = C.c_int(), C.c_int(), C.c_uint64()
    A0, Br, AC, BC = (np.zeros(nproc + 1, np.int32) for _ in range(4))
    lib().orc_part2d(nproc, m, n, k, p(rb), p(rowptr), p(colidx), rA, C.byref(pm), C.byref(pn), C.byref(cost), p(A0), p(Br), p(AC), p(BC))
    return dict(pm=pm.value, pn=pn.value, comm_cost=cost.value, A0_rowptr=A0, B_rowptr=Br[:pm.value + 1].copy(),
                AC_rowptr=AC[:pm.value + 1].copy(), BC_colptr=BC[:pn.value + 1].copy())


def rp_dict(r):
    r = r.contents
    n, P = r.glb_n, r.nproc
    nnz = int(r.A_rowptr[r.A_nrow])
    ns = int(r.rB_sdispls[P]) // n if n else 0
    nr = int(r.rB_rdispls[P]) // n if n else 0
    d = {k: int(getattr(r, k)) for k in ("nproc", "my_rank", "glb_n", "A_nrow", "rB_nrow", "rB_self_src_offset", "rB_self_dst_offset",
                                         "rB_self_nrow", "rB_recv_size")}
    d.update(A_rowptr=arr(r.A_rowptr, r.A_nrow + 1), A_colidx=arr(r.A_colidx, nnz), A_val=arr(r.A_val, nnz, np.float64),
             rB_self_src_ridxs=arr(r.rB_self_src_ridxs, r.rB_self_nrow), rB_scnts=arr(r.rB_scnts, P), rB_sdispls=arr(r.rB_sdispls, P + 1),
             rB_sridxs=arr(r.rB_sridxs, ns), rB_rcnts=arr(r.rB_rcnts, P), rB_rdispls=arr(r.rB_rdispls, P + 1), rB_rridxs=arr(r.rB_rridxs, nr))
    return d


class Simulation:
    """All ranks of one reference run (mode '2d' or 'rp'), simulated in this process by the oracle."""

    def __init__(self, m, k, rowptr, colidx, val, n, mode, nproc, layout=0, reidx=1):
        from pycrp import gen
        L = lib()
        rowptr, colidx, val = i32(rowptr), i32(colidx), np.ascontiguousarray(val, np.float64)
        self.n, self.mode, self.nproc, self.layout = n, mode, nproc, layout
        self.rb = row_partition(rowptr, nproc)
        if mode == "2d":
            self.part = part2d(nproc, m, n, k, self.rb, rowptr, colidx)
            pm, pn = self.part["pm"], self.part["pn"]
            A0, Br, AC, BC = (self.part[x] for x in ("A0_rowptr", "B_rowptr", "AC_rowptr", "BC_colptr"))
        else:
            pm, pn = nproc, 1
            self.part = dict(pm=pm, pn=pn, comm_cost=0)
            A0 = AC = self.rb
            Br = self.rb if m == k else block_split(k, nproc)
            BC = np.array([0, n], np.int32)
            self.part["x_displs"] = Br
        self.pm, self.pn = pm, pn
        # per-rank slices with global nnz offsets in rowptr (examples/test_utils.c:78-91)
        sl = []
        for r in range(nproc):
            s, e = int(A0[r]), int(A0[r + 1])
            sl.append((i32(rowptr[s:e + 1]), i32(colidx[rowptr[s]:rowptr[e]]), np.ascontiguousarray(val[rowptr[s]:rowptr[e]])))
        self.rp = [None] * nproc
        self.keep = []
        for pj in range(pn):
            col_ranks = [pi * pn + pj for pi in range(pm)]
            loc_n = int(BC[pj + 1] - BC[pj])
            plans = (C.POINTER(OrcRp) * pm)()
            for pi in range(pm):
                if mode == "2d":
                    nrow = int(A0[(pi + 1) * pn] - A0[pi * pn])
                    nnz = sum(len(sl[pi * pn + j][1]) for j in range(pn))
                    prp, pci, pv = np.zeros(nrow + 1, np.int32), np.zeros(max(nnz, 1), np.int32), np.zeros(max(nnz, 1), np.float64)
                    rps = (C.c_void_p * pn)(*[p(sl[pi * pn + j][0]) for j in range(pn)])
                    cis = (C.c_void_p * pn)(*[p(sl[pi * pn + j][1]) for j in range(pn)])
                    vs = (C.c_void_p * pn)(*[p(sl[pi * pn + j][2]) for j in range(pn)])
                    L.orc_para2d_panel(pn, pi, p(i32(A0)), rps, cis, vs, p(prp), p(pci), p(pv))
                else:
                    prp, pci, pv = sl[pi]
                    nrow = len(prp) - 1
                self.keep.append((prp, pci, pv))
                plans[pi] = L.orc_rp_create(pm, pi, nrow, p(prp), p(pci), p(pv), p(i32(Br)), loc_n, reidx)
            L.orc_rp_link(plans, pm, p(i32(Br)))
            for pi in range(pm):
                self.rp[col_ranks[pi]] = plans[pi]
        self.Br, self.AC, self.BC, self.A0 = Br, AC, BC, A0
        self.rA_cost = int(L.orc_para2d_rA_cost(int(rowptr[m]), pn))
        self.gen = gen

    def plan(self, rank):
        d = rp_dict(self.rp[rank])
        d.update(pm=self.pm, pn=self.pn, rb_displs0=self.rb)
        for key in ("comm_cost", "A0_rowptr", "B_rowptr", "AC_rowptr", "BC_colptr", "x_displs"):
            if key in self.part:
                d[key] = self.part[key]
        if self.mode == "2d":
            d["rA_cost"] = self.rA_cost
        return d

    def exec(self, dtype=np.float64):
        """C block of every rank (row-major arrays), B = 0.19 i + 0.24 j as the drivers fill it."""
        L = lib()
        pm, pn, lay = self.pm, self.pn, self.layout
        out = [None] * self.nproc
        for pj in range(pn):
            plans = (C.POINTER(OrcRp) * pm)(*[self.rp[pi * pn + pj] for pi in range(pm)])
            sc, nc = int(self.BC[pj]), int(self.BC[pj + 1] - self.BC[pj])
            Bs, Cs, ldB, ldC = [], [], [], []
            for pi in range(pm):
                b0, bn = int(self.Br[pi]), int(self.Br[pi + 1] - self.Br[pi])
                cn = int(self.AC[pi + 1] - self.AC[pi])
                B = self.gen.fill_B(b0, bn, sc, nc)
                if dtype == np.float32:
                    B = B.astype(np.float32).astype(np.float64)     # same inputs the fp32 path sees
                Bs.append(np.ascontiguousarray(B if lay == 0 else B.T))
                Cs.append(np.zeros((cn, nc) if lay == 0 else (nc, cn)))
                ldB.append(nc if lay == 0 else bn)
                ldC.append(nc if lay == 0 else cn)
            Bp = (C.c_void_p * pm)(*[p(b) for b in Bs])
            Cp = (C.c_void_p * pm)(*[p(c) for c in Cs])
            L.orc_rp_exec(plans, pm, lay, Bp, p(i32(ldB)), Cp, p(i32(ldC)))
            for pi in range(pm):
                out[pi * pn + pj] = Cs[pi] if lay == 0 else np.ascontiguousarray(Cs[pi].T)
        return out

    def close(self):
        for r in self.rp:
            lib().orc_rp_free(r)
        self.rp = []
