"""mat_redist_engine (include/mat_redist.h) against the reference's golden dumps: the plan must be
bit-exact, the redistributed blocks identical (pure data movement)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from util import MINIMPIRUN, PKG, run_cmd

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PLAN_KEYS = ("n_proc_send", "n_proc_recv", "send_cnt", "recv_cnt", "send_ranks", "send_sizes", "send_displs", "sblk_sizes",
             "recv_ranks", "recv_sizes", "recv_displs", "rblk_sizes")


def run_redist(tmp_path, name, extra=()):
    lay = cases.redist_layout(name)
    gr, gc = cases.REDIST_DIMS[name]
    path = os.path.join(str(tmp_path), "layout.txt")
    with open(path, "w") as f:
        f.write(f"{len(lay)} {gr} {gc}\n")
        for row in lay:
            f.write(" ".join(str(x) for x in row) + "\n")
    prefix = os.path.join(str(tmp_path), "rd")
    env = dict(os.environ, PYTHONPATH=PKG)
    cmd = [MINIMPIRUN, "-np", str(len(lay)), sys.executable, "-m", "pycrp.redist_flow", path, prefix, *extra]
    r = run_cmd(cmd, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return [dict(np.load(f"{prefix}.r{i}.npz")) for i in range(len(lay))]


def check(name, dumps, f32=False):
    g = dict(np.load(os.path.join(GOLD, "redist_" + name + ".npz")))
    for r, d in enumerate(dumps):
        for key in PLAN_KEYS:
            assert np.array_equal(np.atleast_1d(d[key]), g[f"r{r}/{key}"]), (key, r)
        ref = g[f"r{r}/dst"]
        if f32:
            ref = ref.astype(np.float32)
        assert np.array_equal(d["dst"], ref), r


@pytest.mark.parametrize("name", cases.REDIST_CASES)
def test_redist_host(name, tmp_path):
    check(name, run_redist(tmp_path, name))


@pytest.mark.gpu
@pytest.mark.parametrize("name", cases.REDIST_CASES)
def test_redist_cuda(name, tmp_path):
    check(name, run_redist(tmp_path, name, ("--cuda",)))


@pytest.mark.gpu
def test_redist_cuda_f32(tmp_path):
    check("2x3_to_3x2", run_redist(tmp_path, "2x3_to_3x2", ("--cuda", "--f32")), f32=True)
