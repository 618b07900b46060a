"""Random matrices, random rank counts: the library's host plan on P mini-MPI ranks (no GPU, CRP_SPMM_PLAN_ONLY=1) against the
oracle's single-process simulation of the same reference run - every integer array must be identical.  Complements the fixed
golden cases (tests/test_host_plan.py) with shapes nobody hand-picked."""
import os

import numpy as np
import pytest

import oracle_lib as O
from pycrp import gen
from util import run_flow

CASES = [  # seed, m, k, nnz_per_row, n, mode, nproc, reidx
    (1, 257, 257, 3, 7, "2d", 5, 1),
    (2, 640, 640, 11, 96, "2d", 8, 1),
    (3, 100, 333, 4, 10, "rp", 7, 1),
    (4, 333, 100, 6, 33, "2d", 4, 1),
    (5, 512, 512, 2, 256, "2d", 8, 0),
    (6, 90, 90, 20, 5, "rp", 2, 0),
]


@pytest.mark.parametrize("seed,m,k,npr,n,mode,nproc,reidx", CASES)
def test_random_plan_matches_oracle(seed, m, k, npr, n, mode, nproc, reidx, tmp_path):
    rng = np.random.default_rng(seed)
    empty = tuple(int(x) for x in rng.choice(m - 1, 3, replace=False))      # never the last row: the reference's partition quirk loses it
    mm, kk, rp, ci, v = gen.random_rect(m, k, npr, seed=seed, empty_rows=empty)
    csr = os.path.join(str(tmp_path), "a.bin")
    gen.write_csr_bin(csr, mm, kk, rp, ci, v)
    dumps = run_flow(tmp_path, csr, n, mode, nproc, 0, reidx, plan_only=True)
    sim = O.Simulation(mm, kk, rp, ci, v, n, mode, nproc, 0, reidx)
    for r in range(nproc):
        ref = sim.plan(r)
        for key, val in ref.items():
            if (key == "comm_cost" and mode == "rp") or (key == "rA_cost" and r != 0):      # rA_cost lives on rank 0 only
                continue
            mine = np.atleast_1d(dumps[r][key])
            val = np.atleast_1d(np.asarray(val))
            assert mine.shape == val.shape, (key, r)
            assert np.array_equal(mine.astype(np.float64), val.astype(np.float64)), (key, r)
    sim.close()
