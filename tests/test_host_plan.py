"""The library's host planner (spmat_part.c, rowpara_spmm.c, para2d_spmm.c) on P mini-MPI ranks,
without a GPU (CRP_SPMM_PLAN_ONLY=1), against the golden dumps of the reference: grids, splits,
every index list and count must be bit-exact (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

import cases
from pycrp import gen
from util import run_flow

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SKIP = ("C", "layout", "ldC", "C_nrow", "C_ncol")


@pytest.mark.parametrize("case", cases.SPMM_CASES, ids=[c[0] for c in cases.SPMM_CASES])
def test_plan_matches_reference(case, tmp_path):
    name, spec, n, mode, nproc, layout, reidx = case
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    csr = os.path.join(str(tmp_path), "a.bin")
    gen.write_csr_bin(csr, int(g["m"]), int(g["k"]), g["csr_rowptr"], g["csr_colidx"], g["csr_val"])
    dumps = run_flow(tmp_path, csr, n, mode, nproc, layout, reidx, plan_only=True)
    for r in range(nproc):
        keys = [key[len(f"r{r}/"):] for key in g if key.startswith(f"r{r}/")]
        for key in keys:
            if key in SKIP or (key == "comm_cost" and (r != 0 or mode == "rp")) or (key == "rA_cost" and r != 0):
                continue
            ref = g[f"r{r}/{key}"]
            mine = np.atleast_1d(dumps[r][key])
            assert mine.shape == ref.shape, (key, r, mine.shape, ref.shape)
            assert np.array_equal(mine.astype(np.float64), ref.astype(np.float64)), (key, r)
