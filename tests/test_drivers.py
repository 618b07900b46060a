"""Drop-in check: the reference's own drivers (examples/test_para2d_spmm.c, test_rp_spmm.c, test_spmm_2dpg.c), compiled
UNCHANGED against include/ + libcrpspmm.so by `make drivers`, run on a Matrix-Market file and print the reference's own
figure ||C_ref - C||_f / ||C_ref||_f (examples/test_para2d_spmm.c:212-215) - C_ref from the drivers' single-process CSR loop."""
import os
import re
import subprocess

import numpy as np
import pytest

from pycrp import gen
from util import MINIMPIRUN, PKG, run_cmd

BIN = os.path.join(PKG, "bin")
ROOT = os.path.dirname(PKG)
REF_BIN = os.path.join(ROOT, "oracle", "_ref")


def have(name):
    return os.path.exists(os.path.join(BIN, name))


def write_mtx(tmp_path):
    m, k, rp, ci, v = gen.pwtk_like(m=1500, target_nnz=77000, bandwidth=1200, grid_w=10, seed=5)
    path = os.path.join(str(tmp_path), "small.mtx")
    gen.write_mtx(path, m, k, rp, ci, v)
    return path


@pytest.mark.gpu
@pytest.mark.parametrize("exe,nproc,n", [("test_para2d_spmm.exe", 4, 64), ("test_para2d_spmm.exe", 1, 256), ("test_rp_spmm.exe", 3, 20)])
def test_reference_driver_runs_unchanged(exe, nproc, n, tmp_path):
    if not have(exe):
        pytest.skip("drivers not built (needs the reference sources at build time)")
    mtx = write_mtx(tmp_path)
    r = run_cmd([MINIMPIRUN, "-np", str(nproc), os.path.join(BIN, exe), mtx, str(n), "3", "0", "1"], timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    mobj = re.search(r"\|\|C_ref - C\|\|_f / \|\|C_ref\|\|_f = ([0-9.eE+-]+)", r.stdout)
    assert mobj, r.stdout
    assert float(mobj.group(1)) <= 1e-12, r.stdout
    assert "Local SpMM" in r.stdout and "Redistribute B matrix" in r.stdout          # the reference's stat table rows


def test_partition_driver_matches_reference_build(tmp_path):
    """test_spmm_2dpg (serial, no GPU needed): same grid / cost / splits printed by our build and by the reference build."""
    if not have("test_spmm_2dpg.exe") or not os.path.exists(os.path.join(REF_BIN, "test_spmm_2dpg.exe")):
        pytest.skip("drivers not built")
    mtx = write_mtx(tmp_path)
    outs = []
    for exe in (os.path.join(BIN, "test_spmm_2dpg.exe"), os.path.join(REF_BIN, "test_spmm_2dpg.exe")):
        r = subprocess.run([exe, mtx, "128", "8", "0"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        keep = [ln for ln in r.stdout.splitlines() if "time" not in ln and "Time" not in ln and " used " not in ln and not ln.startswith("Step")]
        outs.append(keep)
    assert outs[0] == outs[1]
    assert any("Calculated 2D grid" in ln for ln in outs[0])
