#!/usr/bin/env python
"""Summarise CRP_PANEL_TRACE files (one per process, written when the SpMM plan is destroyed: 8 time stamps per thread block of
the LAST panel-kernel launch, ns relative to the earliest block start).   python tools/trace_summary.py gpurun_out/prefix.*
The hooks exist only in a library built with  make CRP_NVCC_EXTRA=-DCRP_PANEL_TRACE_BUILD  (they cost ~2 % of the headline kernel)."""
import sys

import numpy as np

COLS = ["start", "put_done", "first_issue", "first_landed", "flag_wait_ns", "last_seen", "first_wait_at", "stop_issued"]
for path in sys.argv[1:]:
    rows = [[int(x) for x in line.split()[1:]] for line in open(path) if not line.startswith("#")]
    if not rows:
        print(path, "empty")
        continue
    t = np.array(rows, dtype=np.float64) / 1e3        # us
    print(f"== {path}: {len(rows)} blocks")
    for j, name in enumerate(COLS):
        v = t[:, j]
        nz = v[v > 0] if name in ("flag_wait_ns", "first_wait_at") else v
        if nz.size == 0:
            print(f"  {name:14s} -")
            continue
        print(f"  {name:14s} min {nz.min():8.2f}  median {np.median(nz):8.2f}  max {nz.max():8.2f} us   (blocks with a value: {nz.size})")
    print(f"  kernel span (earliest start -> latest stop_issued / last_seen): {max(t[:, 7].max(), t[:, 5].max()):.2f} us")
