"""ctypes bindings of the C-ABI of libcrpspmm.so (include/*.h) and of the bundled
mini-MPI, for the tests and bench.py.  The library itself is C + CUDA; nothing here
computes - it only marshals pointers.

Struct layouts mirror include/rowpara_spmm.h, include/para2d_spmm.h and
include/mat_redist.h (which keep the reference's field order,
src/rowpara_spmm.h:8-40, src/para2d_spmm.h:6-14, src/mat_redist.h:7-45).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.environ.get("CRP_LIB_DIR") or os.path.join(PKG_DIR, "lib")      # CRP_LIB_DIR: alternative build (sanitizer runs)
BIN_DIR = os.path.join(PKG_DIR, "bin")

# ---- mini-MPI handles (crp-spmm_b200/minimpi/mpi.h) ----
MPI_COMM_WORLD = 0
MPI_COMM_SELF = 1


def _dt(kind, size):
    return (kind << 8) | size


MPI_BYTE = _dt(2, 1)
MPI_INT = _dt(3, 4)
MPI_UNSIGNED_LONG_LONG = _dt(6, 8)
MPI_FLOAT = _dt(7, 4)
MPI_DOUBLE = _dt(8, 8)
MPI_SUM, MPI_MAX, MPI_MIN = 1, 2, 3

DEV_TYPE_HOST, DEV_TYPE_CUDA, DEV_TYPE_CUDA_MPI_DIRECT = 0, 1, 2

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)


class RowparaSpmm(C.Structure):
    _fields_ = [
        ("nproc", C.c_int), ("my_rank", C.c_int), ("glb_n", C.c_int), ("A_nrow", C.c_int), ("rB_nrow", C.c_int),
        ("rB_self_src_offset", C.c_int), ("rB_self_dst_offset", C.c_int), ("rB_self_nrow", C.c_int),
        ("rB_p2p", C.c_int), ("rB_reidx", C.c_int),
        ("A_rowptr", c_int_p), ("A_colidx", c_int_p), ("rB_self_src_ridxs", c_int_p),
        ("rB_scnts", c_int_p), ("rB_sridxs", c_int_p), ("rB_sdispls", c_int_p),
        ("rB_rcnts", c_int_p), ("rB_rridxs", c_int_p), ("rB_rdispls", c_int_p),
        ("A_val", c_double_p), ("comm", C.c_int),
        ("rB_recv_size", C.c_size_t), ("n_exec", C.c_int),
        ("t_init", C.c_double), ("t_pack", C.c_double), ("t_a2a", C.c_double), ("t_unpack", C.c_double),
        ("t_spmm", C.c_double), ("t_exec", C.c_double),
        ("dev", C.c_void_p),
    ]


class Para2dSpmm(C.Structure):
    _fields_ = [
        ("rp_spmm", C.POINTER(RowparaSpmm)), ("comm_glb", C.c_int), ("comm_col", C.c_int),
        ("rA_cost", C.c_size_t), ("t_init", C.c_double), ("t_ag_A", C.c_double),
        ("t_ag_A_dev", C.c_double), ("ag_A_recv_bytes", C.c_size_t),
    ]


class MatRedistEngine(C.Structure):
    _fields_ = [
        ("graph_comm", C.c_int), ("dtype", C.c_int), ("dt_size", C.c_size_t), ("nproc", C.c_int), ("rank", C.c_int),
        ("src_srow", C.c_int), ("src_scol", C.c_int), ("src_nrow", C.c_int), ("src_ncol", C.c_int),
        ("req_srow", C.c_int), ("req_scol", C.c_int), ("req_nrow", C.c_int), ("req_ncol", C.c_int),
        ("n_proc_send", C.c_int), ("n_proc_recv", C.c_int), ("send_cnt", C.c_int), ("recv_cnt", C.c_int),
        ("alloc_workbuf", C.c_int),
        ("send_ranks", c_int_p), ("send_sizes", c_int_p), ("send_displs", c_int_p), ("sblk_sizes", c_int_p),
        ("recv_ranks", c_int_p), ("recv_sizes", c_int_p), ("recv_displs", c_int_p), ("rblk_sizes", c_int_p),
        ("send_info0", c_int_p), ("recv_info0", c_int_p),
        ("sendbuf_h", C.c_void_p), ("recvbuf_h", C.c_void_p), ("sendbuf_d", C.c_void_p), ("recvbuf_d", C.c_void_p),
        ("workbuf_h", C.c_void_p), ("workbuf_d", C.c_void_p),
        ("hd_trans_ms", C.c_double), ("dev_type", C.c_int), ("dev", C.c_void_p),
    ]


class CrpspmmEngine(C.Structure):
    """struct crpspmm_engine of include/crpspmm.h (public part)."""
    _fields_ = [
        ("np_glb", C.c_int), ("rank_glb", C.c_int), ("np_row", C.c_int), ("np_col", C.c_int), ("rank_row", C.c_int), ("rank_col", C.c_int),
        ("glb_m", C.c_int), ("glb_n", C.c_int), ("glb_k", C.c_int),
        ("loc_A_srow", C.c_int), ("loc_A_erow", C.c_int), ("loc_A_nrow", C.c_int), ("loc_A_nnz", C.c_int),
        ("loc_B_srow", C.c_int), ("loc_B_erow", C.c_int), ("loc_B_scol", C.c_int), ("loc_B_ecol", C.c_int),
        ("loc_B_nrow", C.c_int), ("loc_B_ncol", C.c_int), ("loc_C_srow", C.c_int), ("loc_C_nrow", C.c_int),
        ("alloc_workbuf", C.c_int), ("use_CUDA", C.c_int),
        ("loc_A_rowptr", c_int_p), ("loc_A_colidx", c_int_p), ("loc_A_val", c_double_p),
        ("comm_glb", C.c_int), ("rd_B", C.POINTER(MatRedistEngine)), ("rd_C", C.POINTER(MatRedistEngine)), ("p2d", C.POINTER(Para2dSpmm)),
        ("n_exec", C.c_int), ("t_init", C.c_double), ("t_exec", C.c_double), ("t_rd_A", C.c_double), ("t_agv_A", C.c_double),
        ("t_rd_B", C.c_double), ("t_a2a_B", C.c_double), ("t_spmm", C.c_double), ("t_rd_C", C.c_double), ("t_exec_nr", C.c_double),
        ("nelem_A_rd", C.c_size_t), ("nelem_A_agv", C.c_size_t), ("nelem_B_rd", C.c_size_t), ("nelem_B_a2av", C.c_size_t),
        ("nelem_B_a2av_min", C.c_size_t), ("priv", C.c_void_p),
    ]


# every symbol include/*.h declares (tests/test_cabi_symbols.py checks the library exports them all)
EXPORTS = {
    "utils.h": ["get_wtime_sec", "calc_block_spos_size", "malloc_aligned", "free_aligned", "calc_2norm", "calc_err_2norm",
                "copy_matrix", "print_matrix", "dump_binary"],
    "dev_type.h": ["is_dev_type_valid", "dev_type_malloc", "dev_type_free", "dev_type_realloc", "dev_type_memset",
                   "dev_type_memcpy", "dev_type_copy_matrix"],
    "spmat_part.h": ["csr_mat_row_partition", "csr_mat_row_part_comm_size", "prime_factorization", "calc_spmm_part2d_from_1d"],
    "rowpara_spmm.h": ["rp_spmm_init", "rp_spmm_free", "rp_spmm_exec", "rp_spmm_print_stat", "rp_spmm_clear_stat"],
    "para2d_spmm.h": ["para2d_spmm_init", "para2d_spmm_free", "para2d_spmm_exec", "para2d_spmm_print_stat", "para2d_spmm_clear_stat"],
    "mat_redist.h": ["mat_redist_engine_init", "mat_redist_engine_attach_workbuf", "mat_redist_engine_exec", "mat_redist_engine_free"],
    "crpspmm.h": ["crpspmm_engine_init", "crpspmm_engine_attach_workbuf", "crpspmm_engine_exec", "crpspmm_engine_free",
                  "crpspmm_engine_print_stat", "crpspmm_engine_clear_stat"],
}

_lib = None
_mpi = None


def lib_path():
    return os.path.join(LIB_DIR, "libcrpspmm.so")


def load():
    """Load libminimpi.so (global) and libcrpspmm.so; fails loudly if they were not built."""
    global _lib, _mpi
    if _lib is not None:
        return _lib
    mpi_path = os.path.join(LIB_DIR, "libminimpi.so")
    if not os.path.exists(lib_path()) or not os.path.exists(mpi_path):
        raise RuntimeError(f"{lib_path()} is missing: build it with `make lib` (or __graft_entry__.build()); there is no fallback")
    _mpi = C.CDLL(mpi_path, mode=C.RTLD_GLOBAL)
    _lib = C.CDLL(lib_path(), mode=C.RTLD_GLOBAL)
    L = _lib
    vp, i, d, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    L.get_wtime_sec.restype = d
    L.calc_block_spos_size.argtypes = [i, i, i, c_int_p, c_int_p]
    L.csr_mat_row_partition.argtypes = [i, vp, i, vp]
    L.csr_mat_row_part_comm_size.argtypes = [i, i, vp, vp, i, vp, vp, vp, c_int_p]
    L.prime_factorization.argtypes = [i, C.POINTER(c_int_p)]
    L.prime_factorization.restype = i
    L.calc_spmm_part2d_from_1d.argtypes = [i, i, i, i, vp, vp, vp, i, c_int_p, c_int_p, C.POINTER(sz),
                                           C.POINTER(c_int_p), C.POINTER(c_int_p), C.POINTER(c_int_p), C.POINTER(c_int_p), i]
    L.rp_spmm_init.argtypes = [i, i, vp, vp, vp, vp, i, i, C.POINTER(C.POINTER(RowparaSpmm))]
    L.rp_spmm_free.argtypes = [C.POINTER(C.POINTER(RowparaSpmm))]
    L.rp_spmm_exec.argtypes = [C.POINTER(RowparaSpmm), i, vp, i, vp, i]
    L.rp_spmm_exec_f32.argtypes = [C.POINTER(RowparaSpmm), i, vp, i, vp, i]
    L.rp_spmm_print_stat.argtypes = [C.POINTER(RowparaSpmm)]
    L.rp_spmm_clear_stat.argtypes = [C.POINTER(RowparaSpmm)]
    L.rp_spmm_kernel_name.argtypes = [C.POINTER(RowparaSpmm)]
    L.rp_spmm_kernel_name.restype = C.c_char_p
    L.rp_spmm_set_kernel.argtypes = [C.POINTER(RowparaSpmm), C.c_char_p]
    L.rp_spmm_transport_name.argtypes = [C.POINTER(RowparaSpmm)]
    L.rp_spmm_transport_name.restype = C.c_char_p
    L.rp_spmm_device_times.argtypes = [C.POINTER(RowparaSpmm), c_double_p, c_double_p]
    L.para2d_spmm_init.argtypes = [i, i, i, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.POINTER(Para2dSpmm))]
    L.para2d_spmm_free.argtypes = [C.POINTER(C.POINTER(Para2dSpmm))]
    L.para2d_spmm_exec.argtypes = [C.POINTER(Para2dSpmm), i, vp, i, vp, i]
    L.para2d_spmm_exec_f32.argtypes = [C.POINTER(Para2dSpmm), i, vp, i, vp, i]
    L.para2d_spmm_print_stat.argtypes = [C.POINTER(Para2dSpmm)]
    L.para2d_spmm_clear_stat.argtypes = [C.POINTER(Para2dSpmm)]
    L.mat_redist_engine_init.argtypes = [i] * 8 + [i, i, sz, i, C.POINTER(C.POINTER(MatRedistEngine)), C.POINTER(sz)]
    L.mat_redist_engine_attach_workbuf.argtypes = [C.POINTER(MatRedistEngine), vp, vp]
    L.mat_redist_engine_exec.argtypes = [C.POINTER(MatRedistEngine), vp, i, vp, i]
    L.mat_redist_engine_free.argtypes = [C.POINTER(C.POINTER(MatRedistEngine))]
    L.crpspmm_engine_init.argtypes = [i, i, i, i, i, vp, vp, i, i, i, i, i, i, i, i, i, i, C.POINTER(C.POINTER(CrpspmmEngine)), C.POINTER(sz)]
    L.crpspmm_engine_exec.argtypes = [C.POINTER(CrpspmmEngine), vp, vp, vp, vp, i, vp, i]
    L.crpspmm_engine_free.argtypes = [C.POINTER(C.POINTER(CrpspmmEngine))]
    L.crpspmm_engine_print_stat.argtypes = [C.POINTER(CrpspmmEngine)]
    L.crpspmm_engine_clear_stat.argtypes = [C.POINTER(CrpspmmEngine)]
    L.crpspmm_engine_redist_A_values.argtypes = [C.POINTER(CrpspmmEngine), vp]
    L.is_dev_type_valid.argtypes = [i]
    L.dev_type_malloc.argtypes = [sz, i]
    L.dev_type_malloc.restype = vp
    L.dev_type_free.argtypes = [vp, i]
    L.dev_type_memcpy.argtypes = [vp, vp, sz, i, i]
    L.dev_type_memset.argtypes = [vp, i, sz, i]
    L.dev_type_copy_matrix.argtypes = [sz, i, i, vp, i, vp, i, i]
    # thin CUDA layer (include/crp_cuda.h)
    L.crp_cuda_device_count.restype = i
    L.crp_cuda_sm_count.restype = i
    L.crp_cuda_malloc_dev.argtypes = [C.POINTER(vp), sz]
    L.crp_cuda_malloc_host.argtypes = [C.POINTER(vp), sz]
    L.crp_cuda_free_dev.argtypes = [vp]
    L.crp_cuda_free_host.argtypes = [vp]
    L.crp_cuda_memset_dev.argtypes = [vp, i, sz]
    L.crp_cuda_memset_async.argtypes = [vp, i, sz, vp]
    L.crp_cuda_memcpy_h2d.argtypes = [vp, vp, sz]
    L.crp_cuda_memcpy_d2h.argtypes = [vp, vp, sz]
    L.crp_cuda_memcpy_d2d.argtypes = [vp, vp, sz]
    L.crp_cuda_memcpy_async.argtypes = [vp, vp, sz, vp]
    L.crp_cuda_stream_create.restype = vp
    L.crp_cuda_stream_create_high_priority.restype = vp
    L.crp_cuda_stream_destroy.argtypes = [vp]
    L.crp_cuda_stream_sync.argtypes = [vp]
    L.crp_cuda_event_create.restype = vp
    L.crp_cuda_event_destroy.argtypes = [vp]
    L.crp_cuda_event_record.argtypes = [vp, vp]
    L.crp_cuda_event_sync.argtypes = [vp]
    L.crp_cuda_event_elapsed_ms.argtypes = [vp, vp]
    L.crp_cuda_event_elapsed_ms.restype = C.c_float
    L.crp_cuda_copy_matrix.argtypes = [sz, i, i, vp, i, vp, i]
    L.crp_cuda_gather_rows.argtypes = [sz, i, i, vp, i, vp, vp, i, vp]
    L.crp_cuda_transpose.argtypes = [sz, i, i, vp, i, vp, i, vp]
    L.crp_cuda_spmm_plan_create.argtypes = [i, i, i, vp, vp, vp, i]
    L.crp_cuda_spmm_plan_create.restype = vp
    L.crp_cuda_spmm_plan_destroy.argtypes = [vp]
    L.crp_cuda_spmm_analyse.argtypes = [i, vp, vp, c_int_p, c_int_p, C.POINTER(C.c_longlong), c_int_p]
    L.crp_cuda_spmm_exec.argtypes = [vp, i, i, d, vp, i, vp, i, d, vp, i, vp]
    L.crp_cuda_spmm_last_kernel.argtypes = [vp]
    L.crp_cuda_spmm_last_kernel.restype = C.c_char_p
    L.crp_cuda_spmm_set_variant.argtypes = [vp, C.c_char_p]
    L.crp_cuda_spmm_set_passes.argtypes = [vp, i]
    L.crp_cuda_spmm_last_passes.argtypes = [vp]
    L.crp_cuda_spmm_last_passes.restype = i
    L.crp_cuda_spmm_model_passes.argtypes = [i, i, vp, vp, i, i, d]
    L.crp_cuda_spmm_model_passes.restype = i
    L.crp_cuda_csr_spmm_host.argtypes = [i, i, i, d, i, vp, vp, vp, vp, i, d, vp, i]
    L.crp_cuda_spmm_plan_info.argtypes = [vp, vp]
    L.crp_cuda_measure_dfma_tflops.restype = d
    L.rp_spmm_sync_stats.argtypes = [C.POINTER(RowparaSpmm)]
    L.rp_spmm_plan_info.argtypes = [C.POINTER(RowparaSpmm), vp]
    L.crp_nccl_world_nranks.restype = i
    L.crp_nccl_version.restype = i
    L.crp_set_stream.argtypes = [vp]
    L.crp_set_blocking.argtypes = [i]
    L.crp_kernel_launch_count.restype = C.c_ulonglong
    L.crp_nccl_group_count.restype = C.c_ulonglong
    L.crp_version.restype = C.c_char_p
    # mini-MPI
    M = _mpi
    M.MPI_Init.argtypes = [vp, vp]
    M.MPI_Comm_rank.argtypes = [i, c_int_p]
    M.MPI_Comm_size.argtypes = [i, c_int_p]
    M.MPI_Barrier.argtypes = [i]
    M.MPI_Bcast.argtypes = [vp, i, i, i, i]
    M.MPI_Allreduce.argtypes = [vp, vp, i, i, i, i]
    M.MPI_Gatherv.argtypes = [vp, i, i, vp, vp, vp, i, i, i]
    M.MPI_Wtime.restype = d
    return _lib


def mpi():
    load()
    return _mpi


def ptr(a):
    """void* of a numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def np_from(p, count, dtype):
    """Copy `count` items out of a C pointer into a fresh numpy array."""
    if count <= 0 or not p:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(count,)).astype(dtype, copy=True)


# ---- small MPI conveniences ----
def mpi_init():
    M = mpi()
    M.MPI_Init(None, None)
    r, s = C.c_int(), C.c_int()
    M.MPI_Comm_rank(MPI_COMM_WORLD, C.byref(r))
    M.MPI_Comm_size(MPI_COMM_WORLD, C.byref(s))
    return r.value, s.value


def mpi_bcast(arr, root=0, comm=MPI_COMM_WORLD):
    mpi().MPI_Bcast(ptr(arr), arr.nbytes, MPI_BYTE, root, comm)
    return arr


def mpi_allreduce_max(x, comm=MPI_COMM_WORLD):
    a = np.array([x], dtype=np.float64)
    b = np.zeros(1, dtype=np.float64)
    mpi().MPI_Allreduce(ptr(a), ptr(b), 1, MPI_DOUBLE, MPI_MAX, comm)
    return float(b[0])


def mpi_allreduce_sum(x, comm=MPI_COMM_WORLD):
    a = np.array([x], dtype=np.float64)
    b = np.zeros(1, dtype=np.float64)
    mpi().MPI_Allreduce(ptr(a), ptr(b), 1, MPI_DOUBLE, MPI_SUM, comm)
    return float(b[0])


def mpi_barrier(comm=MPI_COMM_WORLD):
    mpi().MPI_Barrier(comm)


def mpi_finalize():
    mpi().MPI_Finalize()


# ---- device buffers ----
class DevBuf:
    """A device allocation made through the library's own C-ABI (no torch involved)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.p = C.c_void_p()
        load().crp_cuda_malloc_dev(C.byref(self.p), max(self.nbytes, 1))

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        if a.nbytes:
            load().crp_cuda_memcpy_h2d(ptr(a), b.p, a.nbytes)
        return b

    def to_numpy(self, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        if out.nbytes:
            load().crp_cuda_memcpy_d2h(self.p, ptr(out), out.nbytes)
        return out

    def free(self):
        if self.p:
            load().crp_cuda_free_dev(self.p)
            self.p = C.c_void_p()


def rp_plan_dict(rp):
    """All public plan fields of a struct rowpara_spmm as numpy arrays / ints."""
    r = rp.contents if hasattr(rp, "contents") else rp
    np_, n = r.nproc, r.glb_n
    nnz = int(r.A_rowptr[r.A_nrow]) if r.A_nrow >= 0 else 0
    nsend = int(r.rB_sdispls[np_]) // n if n else 0
    nrecv = int(r.rB_rdispls[np_]) // n if n else 0
    d = {k: int(getattr(r, k)) for k in ("nproc", "my_rank", "glb_n", "A_nrow", "rB_nrow", "rB_self_src_offset",
                                         "rB_self_dst_offset", "rB_self_nrow", "rB_recv_size")}
    d["A_rowptr"] = np_from(r.A_rowptr, r.A_nrow + 1, np.int32)
    d["A_colidx"] = np_from(r.A_colidx, nnz, np.int32)
    d["A_val"] = np_from(r.A_val, nnz, np.float64)
    d["rB_self_src_ridxs"] = np_from(r.rB_self_src_ridxs, r.rB_self_nrow, np.int32)
    d["rB_scnts"] = np_from(r.rB_scnts, np_, np.int32)
    d["rB_sdispls"] = np_from(r.rB_sdispls, np_ + 1, np.int32)
    d["rB_sridxs"] = np_from(r.rB_sridxs, nsend, np.int32)
    d["rB_rcnts"] = np_from(r.rB_rcnts, np_, np.int32)
    d["rB_rdispls"] = np_from(r.rB_rdispls, np_ + 1, np.int32)
    d["rB_rridxs"] = np_from(r.rB_rridxs, nrecv, np.int32)
    return d
