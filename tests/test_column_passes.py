"""Column passes (include/crp_cuda.h: crp_cuda_spmm_set_passes / _model_passes): a product whose live window of B / C rows
does not fit the L2 is made in several passes over column blocks.  CPU: the traffic model's decisions.  GPU: any number of
passes gives the same bits as one pass, for every kernel variant, including widths that do not divide evenly."""
import numpy as np
import pytest

from pycrp import capi, gen


def model(mat, n, es, l2):
    m, k, rp, ci, v = mat
    return capi.load().crp_cuda_spmm_model_passes(m, k, capi.ptr(rp), capi.ptr(ci), n, es, float(l2))


def test_model_decisions():
    st = gen.stencil27(32)                         # plane reuse at distance 32*32/32 = 32 row blocks
    # window of one plane pair at n = 1024 fp32: ~ (2 * 1024 + ...) rows * 4 KB ~ 10 MB
    assert model(st, 1024, 4, 126e6) == 1          # fits the B200's L2: one pass
    assert model(st, 1024, 4, 8e6) > 1             # an L2 of 8 MB would not hold it: column passes
    assert model(st, 64, 4, 8e6) == 1              # narrow operands are never split
    pw = gen.pwtk_like(m=20000, target_nnz=1060000, bandwidth=17000, grid_w=32)
    assert model(pw, 256, 8, 126e6) == 1
    er = gen.erdos_renyi(scale=14, nnz_per_row=16)
    assert model(er, 64, 8, 126e6) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["auto", "rowsplit", "mergepath", "panel"])
@pytest.mark.parametrize("n,dtype", [(256, np.float64), (200, np.float64), (384, np.float32)])
def test_passes_are_bitwise_neutral(variant, n, dtype):
    L = capi.load()
    m, k, rp, ci, v = gen.pwtk_like(m=6000, target_nnz=316000, bandwidth=5000, grid_w=16, seed=3)
    es = np.dtype(dtype).itemsize
    B = gen.fill_B(0, k, 0, n, dtype=dtype)
    plan = L.crp_cuda_spmm_plan_create(m, k, k, capi.ptr(rp), capi.ptr(ci), capi.ptr(v), n)
    L.crp_cuda_spmm_set_variant(plan, variant.encode())
    dB, dC = capi.DevBuf.from_numpy(B), capi.DevBuf(m * n * es)
    outs = {}
    for P in (1, 2, 3, 4):
        L.crp_cuda_spmm_set_passes(plan, P)
        L.crp_cuda_memset_async(dC.p, 0xFF, m * n * es, None)
        L.crp_cuda_spmm_exec(plan, n, es, 1.0, dB.p, n, None, 0, 0.0, dC.p, n, None)
        L.crp_cuda_device_sync()
        outs[P] = dC.to_numpy((m, n), dtype)
        got = L.crp_cuda_spmm_last_passes(plan)
        ns = ((n + P - 1) // P + 63) // 64 * 64
        assert got == (1 if (P == 1 or ns >= n) else (n + ns - 1) // ns), (P, got)
        if got > 1:
            assert f"_x{got}pass" in L.crp_cuda_spmm_last_kernel(plan).decode()
    for P in (2, 3, 4):
        assert np.array_equal(outs[1], outs[P]), P
    assert np.isfinite(outs[1]).all()
    dB.free(); dC.free()
    L.crp_cuda_spmm_plan_destroy(plan)
