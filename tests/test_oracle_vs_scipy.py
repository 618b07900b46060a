"""Independent numeric cross-check of the oracle (SURVEY.md §8c "O3"): the C blocks the simulated ranks produce, assembled into
the global matrix, against scipy.sparse on the whole problem.  Catches errors in the oracle's data motion (who sends which row
where) that a comparison against dumps of the same algorithm could share."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle_lib as O
from pycrp import gen


@pytest.mark.parametrize("mode,nproc,n,layout,reidx", [("2d", 8, 48, 0, 1), ("2d", 6, 20, 1, 1), ("rp", 5, 9, 0, 0), ("rp", 4, 16, 1, 1), ("2d", 1, 7, 0, 1)])
@pytest.mark.parametrize("shape", [(240, 240), (150, 310)])
def test_assembled_C_matches_scipy(mode, nproc, n, layout, reidx, shape):
    m, k = shape
    mm, kk, rp, ci, v = gen.random_rect(m, k, 6, seed=m + nproc, empty_rows=(3, 77))
    sim = O.Simulation(mm, kk, rp, ci, v, n, mode, nproc, layout, reidx)
    Cs = sim.exec()
    C = np.full((mm, n), np.nan)
    for r in range(nproc):
        pi, pj = r // sim.pn, r % sim.pn
        r0, r1 = int(sim.AC[pi]), int(sim.AC[pi + 1])
        c0, c1 = int(sim.BC[pj]), int(sim.BC[pj + 1])
        C[r0:r1, c0:c1] = Cs[r]
    covered = int(sim.AC[sim.pm])          # the reference's row split can drop trailing empty rows
    Cref = sp.csr_matrix((v, ci, rp), shape=(mm, kk)) @ gen.fill_B(0, kk, 0, n)
    assert covered == mm or np.all(np.diff(rp)[covered:] == 0)
    err = np.linalg.norm((C[:covered] - Cref[:covered]).ravel()) / np.linalg.norm(Cref.ravel())
    assert err <= 1e-14
    sim.close()
