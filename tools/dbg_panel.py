import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crp-spmm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from pycrp import gen
from test_gpu_spmm import device_spmm, oracle_spmm
from util import rel_err
m, k, rp, ci, v = gen.pwtk_like(m=6000, target_nnz=316000, bandwidth=5000, grid_w=16, seed=11)
for n in [int(x) for x in (sys.argv[1:] or ["256"])]:
    B = np.random.default_rng(n).uniform(-1, 1, (k, n))
    Cd, kern = device_spmm(m, k, rp, ci, v, B, ld_pad=2)
    print(n, kern, rel_err(Cd, oracle_spmm(m, n, rp, ci, v, B)), flush=True)
