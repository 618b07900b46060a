/*
 * Declaration-only stand-in for METIS 5.1.0's <metis.h>: the reference drivers
 * include examples/metis_mat_part.h unconditionally, but METIS itself is an
 * un-vendored dependency (examples/makefile:4) and <part-method>=1 is out of
 * scope.  metis_stub.c aborts if the partitioner is
 * actually requested.
 */
#ifndef ORACLE_METIS_STUB_H
#define ORACLE_METIS_STUB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t idx_t;
typedef float   real_t;

#define METIS_NOPTIONS 40
#define METIS_OK 1
enum { METIS_OPTION_PTYPE = 0, METIS_OPTION_OBJTYPE = 1 };
enum { METIS_OBJTYPE_CUT = 0, METIS_OBJTYPE_VOL = 1 };

int METIS_SetDefaultOptions(idx_t *options);
int METIS_PartGraphKway(
    idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize,
    idx_t *adjwgt, idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options,
    idx_t *edgecut, idx_t *part
);

#ifdef __cplusplus
}
#endif

#endif
