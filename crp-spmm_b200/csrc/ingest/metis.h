/*
 * metis.h - the part of METIS 5's public interface that the reference's driver front-end uses
 * (examples/metis_mat_part.c:31-113 includes <metis.h> through examples/metis_mat_part.h).  METIS is an
 * un-vendored dependency of the reference (examples/makefile:4); here the two entry points are provided
 * by the native partitioner in graph_part.c (libcrpingest.so).  Types as in a default METIS build:
 * 32-bit idx_t, single-precision real_t.
 */
#ifndef CRPSPMM_METIS_SHIM_H
#define CRPSPMM_METIS_SHIM_H

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
/* k-way partition of the graph (xadj, adjncy) with nvtxs vertices into nparts parts: part[v] in [0, nparts), vertex weight of
 * every part at most ubvec * total / nparts (+ 1).  One balance constraint (ncon == 1); vsize, adjwgt, tpwgts are ignored.
 * objval: communication volume (options[METIS_OPTION_OBJTYPE] == METIS_OBJTYPE_VOL) or edge cut.  Returns METIS_OK. */
int METIS_PartGraphKway(
    idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize,
    idx_t *adjwgt, idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options,
    idx_t *objval, idx_t *part
);

#ifdef __cplusplus
}
#endif

#endif
