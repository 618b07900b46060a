"""mat_redist_engine (host path here, CUDA path on the GPU box) on random layouts: the source blocks are a random pr x pc cut of the
matrix, the wanted blocks are arbitrary rectangles (they may overlap each other, be empty, or span several source blocks).
Expected result = the oracle's restatement (oracle/crp_oracle.c, pinned against the reference's golden dumps)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from util import MINIMPIRUN, PKG, run_cmd


def random_layout(seed):
    rng = np.random.default_rng(seed)
    pr, pc = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    P = pr * pc
    R, Cc = int(rng.integers(pr + 3, 60)), int(rng.integers(pc + 3, 50))
    rcut = np.concatenate([[0], np.sort(rng.choice(np.arange(1, R), pr - 1, replace=False)), [R]]) if pr > 1 else np.array([0, R])
    ccut = np.concatenate([[0], np.sort(rng.choice(np.arange(1, Cc), pc - 1, replace=False)), [Cc]]) if pc > 1 else np.array([0, Cc])
    order = rng.permutation(P)                      # which rank holds which source block
    lay = [None] * P
    for b, r in enumerate(order):
        i, j = b // pc, b % pc
        src = (int(rcut[i]), int(ccut[j]), int(rcut[i + 1] - rcut[i]), int(ccut[j + 1] - ccut[j]))
        if rng.random() < 0.15:
            req = (0, 0, 0, 0)                      # wants nothing
        else:
            r0, c0 = int(rng.integers(0, R)), int(rng.integers(0, Cc))
            req = (r0, c0, int(rng.integers(1, R - r0 + 1)), int(rng.integers(1, Cc - c0 + 1)))
        lay[r] = src + req
    return P, R, Cc, lay


def expected(P, lay):
    L = O.lib()
    arr = np.ascontiguousarray(np.array(lay, dtype=np.int32))
    srcs, dsts, sld, dld = [], [], [], []
    for r in range(P):
        s = arr[r]
        ld_s, ld_d = int(s[3]) + 3, int(s[7]) + 2
        a = np.full((int(s[2]), ld_s), -7.0)
        ii, jj = np.meshgrid(np.arange(s[2]), np.arange(s[3]), indexing="ij")
        a[:, :s[3]] = (s[0] + ii) * 1000.5 + (s[1] + jj)
        srcs.append(np.ascontiguousarray(a)); dsts.append(np.full((int(s[6]), ld_d), -1.0)); sld.append(ld_s); dld.append(ld_d)
    Sp = (C.c_void_p * P)(*[O.p(a) for a in srcs])
    Dp = (C.c_void_p * P)(*[O.p(a) for a in dsts])
    L.orc_redist_exec(P, O.p(arr), 8, Sp, O.p(O.i32(sld)), Dp, O.p(O.i32(dld)))
    return dsts


def run(tmp_path, seed, extra=()):
    P, R, Cc, lay = random_layout(seed)
    path = os.path.join(str(tmp_path), "layout.txt")
    with open(path, "w") as f:
        f.write(f"{P} {R} {Cc}\n")
        for row in lay:
            f.write(" ".join(str(x) for x in row) + "\n")
    prefix = os.path.join(str(tmp_path), "rd")
    r = run_cmd([MINIMPIRUN, "-np", str(P), sys.executable, "-m", "pycrp.redist_flow", path, prefix, *extra], env=dict(os.environ, PYTHONPATH=PKG), timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    exp = expected(P, lay)
    for i in range(P):
        got = np.load(f"{prefix}.r{i}.npz")["dst"]
        assert np.array_equal(got, exp[i].ravel()), (seed, i)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8])
def test_random_redist_host(seed, tmp_path):
    run(tmp_path, seed)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 4, 7])
def test_random_redist_cuda(seed, tmp_path):
    run(tmp_path, seed, ("--cuda",))
