"""mat_redist_engine through the C-ABI on P mini-MPI ranks, on the layout files the golden
generator feeds to the reference (oracle/ref_redist_dump.c): G[i, j] = 1000.5 i + j, padded
leading dimensions.  Used by the tests:  minimpirun -np P python -m pycrp.redist_flow layout.txt prefix [--cuda] [--f32]
"""
import ctypes as C
import sys

import numpy as np

from . import capi


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    path, prefix = argv[0], argv[1]
    cuda, f32 = "--cuda" in argv, "--f32" in argv
    rank, nproc = capi.mpi_init()
    L = capi.load()
    with open(path) as f:
        P, gr, gc = (int(x) for x in f.readline().split())
        rows = [[int(x) for x in f.readline().split()] for _ in range(P)]
    assert P == nproc
    r = rows[rank]
    dt = np.float32 if f32 else np.float64
    src_ld, dst_ld = r[3] + 3, r[7] + 2
    src = np.full((max(r[2], 0), src_ld), -7.0, dtype=dt)
    ii, jj = np.meshgrid(np.arange(r[2]), np.arange(r[3]), indexing="ij")
    src[:, :r[3]] = ((r[0] + ii) * 1000.5 + (r[1] + jj)).astype(dt)
    dst = np.full((max(r[6], 0), dst_ld), -1.0, dtype=dt)
    eng = C.POINTER(capi.MatRedistEngine)()
    dev_type = capi.DEV_TYPE_CUDA if cuda else capi.DEV_TYPE_HOST
    L.mat_redist_engine_init(*r, capi.MPI_COMM_WORLD, capi.MPI_FLOAT if f32 else capi.MPI_DOUBLE, dt().itemsize, dev_type, C.byref(eng), None)
    if cuda:
        dsrc, ddst = capi.DevBuf.from_numpy(src), capi.DevBuf.from_numpy(dst)
        for _ in range(2):      # twice: descriptors are cached after the first call
            L.mat_redist_engine_exec(eng, dsrc.p, src_ld, ddst.p, dst_ld)
        dst = ddst.to_numpy(dst.shape, dt)
    else:
        L.mat_redist_engine_exec(eng, capi.ptr(src), src_ld, capi.ptr(dst), dst_ld)
    e = eng.contents
    out = dict(n_proc_send=e.n_proc_send, n_proc_recv=e.n_proc_recv, send_cnt=e.send_cnt, recv_cnt=e.recv_cnt,
               send_ranks=capi.np_from(e.send_ranks, e.n_proc_send, np.int32), send_sizes=capi.np_from(e.send_sizes, e.n_proc_send, np.int32),
               send_displs=capi.np_from(e.send_displs, e.n_proc_send + 1, np.int32), sblk_sizes=capi.np_from(e.sblk_sizes, 4 * e.n_proc_send, np.int32),
               recv_ranks=capi.np_from(e.recv_ranks, e.n_proc_recv, np.int32), recv_sizes=capi.np_from(e.recv_sizes, e.n_proc_recv, np.int32),
               recv_displs=capi.np_from(e.recv_displs, e.n_proc_recv + 1, np.int32), rblk_sizes=capi.np_from(e.rblk_sizes, 4 * e.n_proc_recv, np.int32),
               dst_ld=dst_ld, dst=dst.ravel())
    np.savez(f"{prefix}.r{rank}.npz", **out)
    L.mat_redist_engine_free(C.byref(eng))
    capi.mpi_barrier()
    capi.mpi_finalize()
    return 0


if __name__ == "__main__":
    sys.exit(main())
