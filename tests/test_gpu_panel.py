"""GPU parity of the B-row-panel kernel (crp-spmm_b200/csrc/cuda/spmm_panel.cu: cp.async.bulk + mbarrier pipeline,
warp-specialised) against the oracle, through the C-ABI.  Replaces mkl_sparse_d_mm at reference
src/rowpara_spmm.c:398-408.  Tolerances: BASELINE.json (1e-12 fp64, 1e-5 fp32)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from pycrp import gen
from test_gpu_spmm import TOL32, TOL64, device_spmm, oracle_spmm
from test_panel_structure import perturb
from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pw():
    return gen.pwtk_like(m=6000, target_nnz=316000, bandwidth=5000, grid_w=16, seed=11)


def with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("n", [64, 66, 128, 200, 256, 258, 512, 640])
def test_panel_fp64(pw, n):
    m, k, rp, ci, v = pw
    B = np.random.default_rng(n).uniform(-1, 1, (k, n))
    Cd, kern = device_spmm(m, k, rp, ci, v, B, ld_pad=2)
    assert "panel_f64_R6" in kern, kern
    assert rel_err(Cd, oracle_spmm(m, n, rp, ci, v, B)) <= TOL64


@pytest.mark.parametrize("n", [64, 256, 520, 1024])
def test_panel_fp32(pw, n):
    m, k, rp, ci, v = pw
    B = np.random.default_rng(n).uniform(-1, 1, (k, n)).astype(np.float32)
    Cd, kern = device_spmm(m, k, rp, ci, v, B, dtype=np.float32, ld_pad=4)
    assert "panel_f32_R6" in kern, kern
    assert rel_err(Cd, oracle_spmm(m, n, rp, ci, v.astype(np.float32).astype(np.float64), B.astype(np.float64))) <= TOL32


@pytest.mark.parametrize("cfg", [dict(CRP_PANEL_CR="4", CRP_PANEL_EMAX="8"), dict(CRP_PANEL_CR="16", CRP_PANEL_STAGES="2"),
                                 dict(CRP_PANEL_CR="48"), dict(CRP_PANEL_GRID="3")])
def test_panel_configurations(pw, cfg):
    """tile width, chunk size, pipeline depth and grid size change the schedule, never the result (bit for bit)"""
    m, k, rp, ci, v = pw
    B = np.random.default_rng(1).uniform(-1, 1, (k, 256))
    base, kern0 = device_spmm(m, k, rp, ci, v, B)
    Cd, kern = with_env(cfg, lambda: device_spmm(m, k, rp, ci, v, B))
    assert "panel" in kern0 and "panel" in kern
    assert np.array_equal(Cd, base)


def test_panel_matches_rowsplit_bitwise(pw):
    """same accumulation order as the CSR row (ascending columns): identical bits to the row-split kernel"""
    m, k, rp, ci, v = pw
    B = np.random.default_rng(2).uniform(-1, 1, (k, 128))
    a, ka = device_spmm(m, k, rp, ci, v, B, variant=b"panel")
    b, kb = device_spmm(m, k, rp, ci, v, B, variant=b"rowsplit")
    assert "panel" in ka and "rowsplit" in kb
    assert np.array_equal(a, b)


def test_panel_alpha_beta_two_piece(pw):
    m, k, rp, ci, v = pw
    rng = np.random.default_rng(3)
    B, C0 = rng.uniform(-1, 1, (k, 128)), rng.uniform(-1, 1, (m, 128))
    Cref = 0.5 * oracle_spmm(m, 128, rp, ci, v, B) - 2.0 * C0
    Cd, kern = device_spmm(m, k, rp, ci, v, B, alpha=0.5, beta=-2.0, C0=C0, x0_rows=1234)
    assert "panel" in kern, kern
    assert rel_err(Cd, Cref) <= TOL64


@pytest.mark.parametrize("frac", [0.01, 0.05])
def test_panel_relaxed_groups(pw, frac):
    """near-identical groups (boundary conditions of a real FEM matrix): masked blocks, absent entries never multiplied"""
    m, k, rp, ci, v = perturb(pw, frac)
    B = np.random.default_rng(4).uniform(-1, 1, (k, 256))
    B[::97] = np.inf                                     # an Inf row of B must only reach the C rows that really reference it
    Cd, kern = device_spmm(m, k, rp, ci, v, B)
    assert "panel_f64_R6" in kern, kern
    b, _ = device_spmm(m, k, rp, ci, v, B, variant=b"rowsplit")
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(Cd), fin)
    assert np.array_equal(Cd[fin], b[fin])


def test_panel_stencil_low_fill():
    m, k, rp, ci, v = gen.stencil27(16)
    B = np.random.default_rng(5).uniform(-1, 1, (k, 128)).astype(np.float32)
    Cd, kern = with_env(dict(CRP_SPMM_RG_FILL="0.3"), lambda: device_spmm(m, k, rp, ci, v, B, dtype=np.float32))
    assert "panel_f32" in kern, kern
    assert rel_err(Cd, oracle_spmm(m, 128, rp, ci, v, B.astype(np.float64))) <= TOL32


def test_misaligned_operands_fall_back(pw):
    """a leading dimension that is not a multiple of 16 bytes cannot be bulk-copied: the register-blocked kernel runs"""
    m, k, rp, ci, v = pw
    B = np.random.default_rng(6).uniform(-1, 1, (k, 64))
    Cd, kern = device_spmm(m, k, rp, ci, v, B, ld_pad=1)
    assert "panel" not in kern, kern
    assert rel_err(Cd, oracle_spmm(m, 64, rp, ci, v, B)) <= TOL64
