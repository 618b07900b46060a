"""bench.py's multi-GPU launch path on CPU: two ranks started by torchrun (as the driver does for N > 1,
rendezvous on 127.0.0.1) must form a mini-MPI world from RANK / WORLD_SIZE / MASTER_PORT and produce the
reference's plan (CRP_SPMM_PLAN_ONLY=1: no GPU here)."""
import os
import subprocess
import sys

import numpy as np

from pycrp import gen
from util import PKG

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_two_ranks_under_torchrun(tmp_path):
    g = dict(np.load(os.path.join(GOLD, "rand300_2d_np2_n16.npz")))
    csr = os.path.join(str(tmp_path), "a.bin")
    gen.write_csr_bin(csr, int(g["m"]), int(g["k"]), g["csr_rowptr"], g["csr_colidx"], g["csr_val"])
    prefix = os.path.join(str(tmp_path), "dump")
    env = dict(os.environ, PYTHONPATH=PKG, CRP_SPMM_PLAN_ONLY="1", OMP_NUM_THREADS="1")
    for k in ("MINIMPI_DIR", "MINIMPI_RANK", "MINIMPI_SIZE"):
        env.pop(k, None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29671", "-m", "pycrp.flow", csr, "16", "2d", "--no-exec", "--dump", prefix]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for rank in range(2):
        d = dict(np.load(f"{prefix}.r{rank}.npz"))
        assert int(d["nproc"]) == int(g[f"r{rank}/nproc"][0])
        for key in ("A0_rowptr", "B_rowptr", "AC_rowptr", "BC_colptr", "A_rowptr", "A_colidx", "rB_sridxs", "rB_rridxs", "rB_scnts", "rB_rcnts"):
            assert np.array_equal(d[key], g[f"r{rank}/{key}"]), (key, rank)
        assert int(d["pm"]) == int(g[f"r{rank}/pm"][0]) and int(d["pn"]) == int(g[f"r{rank}/pn"][0])
