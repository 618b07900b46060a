"""The multi-rank data planes under a parity assertion on ONE GPU (the driver's GPU test box has one): ranks that share
a device run the peer-memory transport through CUDA IPC with a host barrier instead of device-side spinning
("p2p-hostsync": put + signal kernel, receive halves, the SpMM kernel's flag reads and wait map all execute), the
host-staged transport, and the overlap split (own-rows pass, then received-rows pass with beta = 1) - all against the
reference's golden C (tests/golden/, minted from oracle/_ref) at 1e-12.
Replaces the exchange at reference src/rowpara_spmm.c:266-311; check as in examples/test_para2d_spmm.c:170-221."""
import os

import numpy as np
import pytest

import cases
import oracle_lib as O
from pycrp import gen
from util import run_flow

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# one case per feature (2-D grid with pn > 1, column-major, rp flow, m != k, panel kernel on 8 ranks); every case of cases.SPMM_CASES
# runs through the default transport in tests/test_gpu_spmm.py
NAMES = ["tridiag16_2d_np8_n16", "rand300_2d_np4_n16_cm", "rand300_rp_np4_n8", "rect200x350_2d_np4_n12", "pwtk600_2d_np8_n64"]
CASES = [c for c in cases.SPMM_CASES if c[0] in NAMES]


@pytest.mark.parametrize("overlap", ["0", "1"])
@pytest.mark.parametrize("transport", ["1", "2"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_engine_transports_match_reference_golden(case, transport, overlap, tmp_path):
    name, spec, n, mode, nproc, layout, reidx = case
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    csr = os.path.join(str(tmp_path), "a.bin")
    gen.write_csr_bin(csr, int(g["m"]), int(g["k"]), g["csr_rowptr"], g["csr_colidx"], g["csr_val"])
    dumps = run_flow(tmp_path, csr, n, mode, nproc, layout, reidx, device=True, extra_env={"CRP_SPMM_TRANSPORT": transport, "CRP_SPMM_OVERLAP": overlap})
    num = den = 0.0
    for r in range(nproc):
        nrow, ncol = int(g[f"r{r}/C_nrow"][0]), int(g[f"r{r}/C_ncol"][0])
        Cref = g[f"r{r}/C"]
        Cref = Cref.reshape(nrow, ncol) if layout == 0 else Cref.reshape(ncol, nrow).T
        num += float(np.sum((dumps[r]["C"] - Cref) ** 2)); den += float(np.sum(Cref ** 2))
        tr = str(dumps[r]["transport"])
        if int(dumps[r]["nproc"]) > 1:          # ranks of a grid column (pm); pm == 1 -> no exchange at all
            want = {"1": "staged", "2": "p2p-hostsync"}[transport]
            has_comm = int(dumps[r]["rB_recv_size"]) > 0 or int(np.sum(dumps[r]["rB_scnts"])) > 0
            assert tr == want + ("+overlap" if overlap == "1" and has_comm else ""), tr
    assert np.sqrt(num) <= 1e-12 * np.sqrt(den)


@pytest.mark.parametrize("transport,overlap", [("2", "0"), ("2", "1"), ("1", "1")])
def test_engine_transports_fp32_and_repeat(transport, overlap, tmp_path):
    """fp32 entry points over the same data planes; several execs in a row alternate the two receive halves"""
    m, k, rp, ci, v = gen.pwtk_like(m=1800, target_nnz=90000, bandwidth=1500, grid_w=10, seed=3)
    csr = os.path.join(str(tmp_path), "a.bin")
    gen.write_csr_bin(csr, m, k, rp, ci, v)
    sim = O.Simulation(m, k, rp, ci, v.astype(np.float32).astype(np.float64), 64, "rp", 4)
    Cs = sim.exec(dtype=np.float32)
    import subprocess, sys
    from util import MINIMPIRUN, PKG, run_cmd
    prefix = os.path.join(str(tmp_path), "dump")
    env = dict(os.environ, PYTHONPATH=PKG, CRP_SPMM_TRANSPORT=transport, CRP_SPMM_OVERLAP=overlap, OMP_NUM_THREADS="2")
    r = run_cmd([MINIMPIRUN, "-np", "4", sys.executable, "-m", "pycrp.flow", csr, "64", "rp", "--dump", prefix, "--device", "--f32", "--ntest", "5"], env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    num = den = 0.0
    for q in range(4):
        d = dict(np.load(f"{prefix}.r{q}.npz"))
        num += float(np.sum((d["C"].astype(np.float64) - Cs[q]) ** 2)); den += float(np.sum(Cs[q] ** 2))
        assert "panel_f32" in str(d["kernel"]), d["kernel"]
    assert np.sqrt(num) <= 1e-5 * np.sqrt(den)
