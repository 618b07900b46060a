#!/usr/bin/env python
"""Where does the per-step time outside the SpMM kernel go?  (development aid)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crp-spmm_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench
from pycrp import capi
from pycrp.flow import Problem

L = capi.load()
capi.mpi_init()
pb = Problem(bench.matrix_path("pwtk"), 256, "2d").init()
stream = L.crp_cuda_stream_create(); L.crp_set_stream(stream)
B = pb.make_B(); C_ = pb.empty_C()
dB, dC = capi.DevBuf.from_numpy(B), capi.DevBuf(C_.nbytes)
dF = capi.DevBuf(256 << 20)
K = 20
ev = [(L.crp_cuda_event_create(), L.crp_cuda_event_create()) for _ in range(K)]
for mode in ("flush+nonblocking", "noflush+nonblocking", "flush+blocking", "flush+nonblocking+sleep"):
    L.crp_set_blocking(0 if "nonblocking" in mode else 1)
    for it in range(3):
        pb.exec_ptr(dB.p, dC.p)
    L.crp_cuda_stream_sync(stream); pb.clear_stat()
    host = []
    t0 = time.perf_counter()
    for it in range(K):
        if mode.startswith("flush"):
            L.crp_cuda_memset_async(dF.p, it, 256 << 20, stream)
        L.crp_cuda_event_record(ev[it][0], stream)
        h0 = time.perf_counter()
        pb.exec_ptr(dB.p, dC.p)
        host.append(time.perf_counter() - h0)
        L.crp_cuda_event_record(ev[it][1], stream)
        if "sleep" in mode:
            time.sleep(0.002)
    L.crp_cuda_stream_sync(stream)
    wall = (time.perf_counter() - t0) / K
    pair = [L.crp_cuda_event_elapsed_ms(a, b) for a, b in ev]
    gaps = [L.crp_cuda_event_elapsed_ms(ev[i][1], ev[i + 1][0]) for i in range(K - 1)]
    r = pb.rp.contents
    L.rp_spmm_print_stat if False else None
    d0, d1 = capi.C.c_double(), capi.C.c_double(); L.rp_spmm_device_times(pb.rp, capi.C.byref(d0), capi.C.byref(d1))
    print(f"{mode:28s} pair_ms median {np.median(pair):.4f} min {min(pair):.4f} max {max(pair):.4f} | t_spmm {1e3*r.t_spmm/max(r.n_exec,1):.4f} | host exec call us median {1e6*np.median(host):.1f} | wall/step ms {1e3*wall:.4f} | gap ev1->next ev0 ms {np.median(gaps):.4f}")
L.crp_set_blocking(1)
