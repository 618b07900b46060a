/* METIS is not available offline; <part-method>=1 is out of scope (see metis.h). */
#include <stdio.h>
#include <stdlib.h>
#include "metis.h"

int METIS_SetDefaultOptions(idx_t *options)
{
    for (int i = 0; i < METIS_NOPTIONS; i++) options[i] = -1;
    return METIS_OK;
}

int METIS_PartGraphKway(
    idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize,
    idx_t *adjwgt, idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options,
    idx_t *edgecut, idx_t *part
)
{
    (void) nvtxs; (void) ncon; (void) xadj; (void) adjncy; (void) vwgt; (void) vsize; (void) adjwgt;
    (void) nparts; (void) tpwgts; (void) ubvec; (void) options; (void) edgecut; (void) part;
    fprintf(stderr, "METIS is not available in this build: use <part-method> = 0\n");
    abort();
    return 0;
}
