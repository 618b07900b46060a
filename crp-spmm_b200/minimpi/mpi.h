/*
 * minimpi - a single-node, process-per-rank subset of the MPI C API.
 *
 * Why it exists: the CRP-SpMM public headers take an MPI_Comm and the drop-in
 * drivers (test_para2d_spmm, test_rp_spmm) call MPI directly, but neither the
 * build container nor the B200 boxes ship an MPI.  This shim implements exactly
 * the MPI surface the library, the drivers and the oracle build use (SURVEY.md
 * App. B) over unix-domain stream sockets, so that the whole stack is
 * self-contained.  It is the *control plane* only: bulk device data moves with
 * NCCL, never through here.
 *
 * Launch: `minimpirun -np N prog args...`, or under torchrun (RANK / WORLD_SIZE
 * / LOCAL_RANK / MASTER_PORT are honoured), or stand-alone (size 1).
 *
 * If a real MPI is installed, compile against its <mpi.h> instead; nothing in
 * the library depends on minimpi internals.
 */
#ifndef MINIMPI_MPI_H
#define MINIMPI_MPI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MINIMPI 1
#define MPI_VERSION 3
#define MPI_SUBVERSION 1

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef struct minimpi_request *MPI_Request;

typedef struct MPI_Status
{
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
    size_t minimpi_nbytes;
} MPI_Status;

#define MPI_SUCCESS        0
#define MPI_ERR_OTHER      15

#define MPI_COMM_NULL      ((MPI_Comm) -1)
#define MPI_COMM_WORLD     ((MPI_Comm) 0)
#define MPI_COMM_SELF      ((MPI_Comm) 1)

#define MPI_REQUEST_NULL   ((MPI_Request) 0)
#define MPI_STATUS_IGNORE  ((MPI_Status *) 0)
#define MPI_STATUSES_IGNORE ((MPI_Status *) 0)
#define MPI_INFO_NULL      ((MPI_Info) 0)
#define MPI_UNWEIGHTED     ((int *) 2)
#define MPI_IN_PLACE       ((void *) 1)
#define MPI_ANY_SOURCE     (-2)
#define MPI_ANY_TAG        (-1)
#define MPI_UNDEFINED      (-32766)
#define MPI_MAX_PROCESSOR_NAME 256

/* datatype handle = (kind << 8) | size-in-bytes */
#define MINIMPI_DT(kind, size) (((kind) << 8) | (size))
#define MPI_DATATYPE_NULL        0
#define MPI_CHAR                 MINIMPI_DT(1, 1)
#define MPI_BYTE                 MINIMPI_DT(2, 1)
#define MPI_INT                  MINIMPI_DT(3, 4)
#define MPI_UNSIGNED             MINIMPI_DT(4, 4)
#define MPI_LONG_LONG            MINIMPI_DT(5, 8)
#define MPI_LONG_LONG_INT        MPI_LONG_LONG
#define MPI_LONG                 MPI_LONG_LONG
#define MPI_UNSIGNED_LONG_LONG   MINIMPI_DT(6, 8)
#define MPI_UNSIGNED_LONG        MPI_UNSIGNED_LONG_LONG
#define MPI_FLOAT                MINIMPI_DT(7, 4)
#define MPI_DOUBLE               MINIMPI_DT(8, 8)
#define MPI_INT64_T              MPI_LONG_LONG
#define MPI_UINT64_T             MPI_UNSIGNED_LONG_LONG
#define MPI_INT32_T              MPI_INT

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

/* environment */
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Initialized(int *flag);
int MPI_Finalized(int *flag);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);
int MPI_Get_processor_name(char *name, int *resultlen);

/* communicators */
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm);
int MPI_Dims_create(int nnodes, int ndims, int dims[]);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Dist_graph_create_adjacent(
    MPI_Comm comm_old, int indegree, const int *sources, const int *sourceweights,
    int outdegree, const int *destinations, const int *destweights,
    MPI_Info info, int reorder, MPI_Comm *comm_dist_graph
);

/* point to point */
int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status statuses[]);
int MPI_Get_count(const MPI_Status *status, MPI_Datatype dt, int *count);
int MPI_Type_size(MPI_Datatype dt, int *size);

/* collectives */
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Ibcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm, MPI_Request *req);
int MPI_Gather(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Gatherv(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, const int rcounts[], const int displs[], MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Scatterv(const void *sbuf, const int scounts[], const int displs[], MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Iscatterv(const void *sbuf, const int scounts[], const int displs[], MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm, MPI_Request *req);
int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Allgatherv(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, const int rcounts[], const int displs[], MPI_Datatype rdt, MPI_Comm comm);
int MPI_Iallgatherv(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, const int rcounts[], const int displs[], MPI_Datatype rdt, MPI_Comm comm, MPI_Request *req);
int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Alltoallv(const void *sbuf, const int scounts[], const int sdispls[], MPI_Datatype sdt, void *rbuf, const int rcounts[], const int rdispls[], MPI_Datatype rdt, MPI_Comm comm);
int MPI_Neighbor_alltoallv(const void *sbuf, const int scounts[], const int sdispls[], MPI_Datatype sdt, void *rbuf, const int rcounts[], const int rdispls[], MPI_Datatype rdt, MPI_Comm comm);
int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);

/* minimpi extensions (not MPI): world rank of a communicator member, node-local rank */
int minimpi_comm_world_rank(MPI_Comm comm, int rank_in_comm);
int minimpi_local_rank(void);

#ifdef __cplusplus
}
#endif

#endif
