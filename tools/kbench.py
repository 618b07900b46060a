#!/usr/bin/env python
"""Kernel-only timing of the local SpMM variants on one GPU (development aid; bench.py is the contract).
   python tools/kbench.py [--workload pwtk] [--n 256] [--variants auto,rowsplit] [--iters 10] [--env K=V,...]
Each variant is timed with CUDA events around single launches, L2 flushed in between.
Variant specs may carry environment settings, e.g. auto:CRP_PANEL_CLUSTER=0 or panel:CRP_PANEL_CR=24:CRP_PANEL_EMAX=96."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crp-spmm_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from pycrp import capi, gen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="pwtk")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--variants", default="auto,rowsplit")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--f32", action="store_true")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--rows", default="", help="a:b as fractions of m, e.g. 0.125:0.25 - time the kernel on a row slice (what one rank of several holds)")
    ap.add_argument("--passes", default="0", help="comma list of column-pass counts to time per variant (0 = automatic)")
    ap.add_argument("--ldb0", action="store_true", help="experiment: leading dimension 0, every B row aliases row 0 (all gathers hit L1)")
    a = ap.parse_args()
    L = capi.load()
    gname, gkw, n, dts, mode, desc = bench.WORKLOADS[a.workload]
    n = a.n or n
    dt = np.float32 if (a.f32 or dts == "f32") else np.float64
    m, k, rp, ci, v = gen.read_csr_bin(bench.matrix_path(a.workload))
    if a.rows:
        f0, f1 = (float(x) for x in a.rows.split(":"))
        r0, r1 = int(f0 * m), int(f1 * m)
        rp, ci, v = np.ascontiguousarray(rp[r0:r1 + 1] - rp[r0]), np.ascontiguousarray(ci[rp[r0]:rp[r1]]), np.ascontiguousarray(v[rp[r0]:rp[r1]])
        m = r1 - r0
    nnz = int(rp[-1])
    es = np.dtype(dt).itemsize
    B = gen.fill_B(0, k, 0, n, dtype=dt)
    dB, dC = capi.DevBuf.from_numpy(B), capi.DevBuf(m * n * es)
    dF = capi.DevBuf(256 << 20)
    stream = L.crp_cuda_stream_create()
    e0, e1 = L.crp_cuda_event_create(), L.crp_cuda_event_create()
    algo = 4 * (m + 1) + (4 + es) * nnz + es * k * n + es * m * n
    ref = None
    for spec in a.variants.split(","):
        # spec: variant[:ENV=VAL[:ENV=VAL]]
        parts = spec.split(":")
        for kv in parts[1:]:
            key, val = kv.split("=")
            os.environ[key] = val
        plan = L.crp_cuda_spmm_plan_create(m, k, k, capi.ptr(rp), capi.ptr(ci), capi.ptr(v), n)
        L.crp_cuda_spmm_set_variant(plan, parts[0].encode())
        for npass in (int(x) for x in a.passes.split(",")):
            L.crp_cuda_spmm_set_passes(plan, npass)
            ts = []
            for it in range(a.iters + 2):
                L.crp_cuda_memset_async(dF.p, it, 256 << 20, stream)
                L.crp_cuda_event_record(e0, stream)
                L.crp_cuda_spmm_exec(plan, n, es, 1.0, dB.p, 0 if a.ldb0 else n, None, 0, 0.0, dC.p, n, stream)
                L.crp_cuda_event_record(e1, stream)
                L.crp_cuda_event_sync(e1)
                if it >= 2:
                    ts.append(L.crp_cuda_event_elapsed_ms(e0, e1))
            name = L.crp_cuda_spmm_last_kernel(plan).decode()
            ms = float(np.median(ts))
            out = {"spec": spec, "passes": L.crp_cuda_spmm_last_passes(plan), "kernel": name, "ms_median": ms, "ms_min": min(ts), "gflops": 2.0 * nnz * n / ms / 1e6, "algo_gbs": algo / ms / 1e6, "n": n, "dtype": str(np.dtype(dt))}
            if a.check:
                Cd = dC.to_numpy((m, n), dt)
                if ref is None:
                    ref = Cd
                out["max_rel_diff_vs_first"] = float(np.max(np.abs(Cd.astype(np.float64) - ref) / (np.abs(ref) + 1e-300)))
            print(json.dumps(out), flush=True)
        L.crp_cuda_spmm_plan_destroy(plan)
        for kv in parts[1:]:
            os.environ.pop(kv.split("=")[0], None)


if __name__ == "__main__":
    main()
