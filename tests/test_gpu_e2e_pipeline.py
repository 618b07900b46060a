"""Host-resident B / C (the reference's calling convention) through the pipelined path of rp_spmm_exec: the call is cut into
column panels that flow H2D | exchange + product | D2H on three streams (crp-spmm_b200/csrc/host/rowpara_spmm.c,
rp_e2e_panel_count).  Whatever the number of panels, C must be bit-identical to the serial path and to the device-resident
call; the engine must say which route it took.  Multi-rank runs on a box with one GPU share the device and take the serial
route (a panel is an exchange round of the device-spinning transport) - the multi-rank pipeline is exercised by
`bench.py --gpus N` (its e2e leg checks C against the device-resident result, `matches_device_result`)."""
import os

import numpy as np
import pytest

from pycrp import gen
from util import run_flow

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,f32", [(256, False), (200, False), (384, True)])
def test_panels_are_bitwise_neutral(n, f32, tmp_path):
    m, k, rp, ci, v = gen.pwtk_like(m=6000, target_nnz=316000, bandwidth=5000, grid_w=16, seed=3)
    csr = os.path.join(str(tmp_path), "a.bin")
    gen.write_csr_bin(csr, m, k, rp, ci, v)
    ref = run_flow(tmp_path, csr, n, "2d", 1, device=True, f32=f32)[0]["C"]
    for panels in ("1", "4", "7"):
        out = run_flow(tmp_path, csr, n, "2d", 1, device=False, f32=f32, extra_env={"CRP_SPMM_E2E_PANELS": panels})[0]
        assert np.array_equal(out["C"], ref), panels
    assert np.isfinite(ref).all() and np.abs(ref).max() > 0


@pytest.mark.parametrize("nproc", [2, 4])
def test_multi_rank_host_buffers_still_exact(nproc, tmp_path):
    """ranks sharing the GPU: serial route for every rank (the decision is collective), results as with device buffers"""
    m, k, rp, ci, v = gen.pwtk_like(m=3000, target_nnz=150000, bandwidth=2500, grid_w=12, seed=5)
    csr = os.path.join(str(tmp_path), "a.bin")
    gen.write_csr_bin(csr, m, k, rp, ci, v)
    dev = run_flow(tmp_path, csr, 128, "rp", nproc, device=True)
    host = run_flow(tmp_path, csr, 128, "rp", nproc, device=False, extra_env={"CRP_SPMM_E2E_PANELS": "4"})
    for r in range(nproc):
        assert np.array_equal(dev[r]["C"], host[r]["C"]), r
