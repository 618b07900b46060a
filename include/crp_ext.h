/*
 * crp_ext.h - additions of the B200 build that have no counterpart in the
 * reference API.  Existing callers never need them; bench.py and the fp32
 * configuration (BASELINE.json config 5) do.
 */
#ifndef CRPSPMM_CRP_EXT_H
#define CRPSPMM_CRP_EXT_H

#include "rowpara_spmm.h"
#include "para2d_spmm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* fp32 B and C (A's values are rounded to fp32 once, on the device).  Same rules as the fp64 calls. */
void rp_spmm_exec_f32(rp_spmm_p rp_spmm, const int BC_layout, const float *B, const int ldB, float *C, const int ldC);
void para2d_spmm_exec_f32(para2d_spmm_p para2d_spmm, const int BC_layout, const float *B, const int ldB, float *C, const int ldC);

/* By default every engine launches on its own non-blocking stream and exec
 * returns after the result is complete (reference semantics).  A caller that
 * manages streams itself can (a) make all engines created afterwards launch on
 * `stream` (a cudaStream_t; NULL restores the default) and (b) drop the final
 * synchronisation so that exec only enqueues work. */
void crp_set_stream(void *stream);
void crp_set_blocking(const int blocking);

/* Number of kernels launched / NCCL groups issued by this process since start. */
unsigned long long crp_kernel_launch_count(void);
unsigned long long crp_nccl_group_count(void);

/* Size of the NCCL communicator spanning MPI_COMM_WORLD (created on first use; collective), NCCL's version code. */
int crp_nccl_world_nranks(void);
int crp_nccl_version(void);

/* Name of the local-SpMM kernel variant the engine's last exec launched. */
const char *rp_spmm_kernel_name(rp_spmm_p rp_spmm);
/* Data plane of the B-row exchange this engine uses: "single" (one rank), "p2p" (NVLink peer stores + arrival flags, the
 * SpMM kernel waits itself), "p2p-hostsync" (same stores between ranks that share a GPU, arrival by host barrier), "nccl",
 * "staged" (pinned host memory + MPI); "+overlap" when the product is split into own-rows / received-rows passes. */
const char *rp_spmm_transport_name(rp_spmm_p rp_spmm);
/* Force a variant ("auto", "rowsplit", "rowgroup", "panel", "mergepath") for experiments. */
void rp_spmm_set_kernel(rp_spmm_p rp_spmm, const char *name);

/* Device time (seconds since the last clear_stat) spent staging a host B in / a host C out. */
void rp_spmm_device_times(rp_spmm_p rp_spmm, double *t_h2d, double *t_d2h);

/* Shape of the local-SpMM plan (crp_cuda_spmm_plan_info of include/crp_cuda.h) of this engine. */
void rp_spmm_plan_info(rp_spmm_p rp_spmm, long long out[12]);

/* Fold the CUDA-event times of every exec issued so far into the engine's public counters (t_pack, t_a2a, t_spmm,
 * t_exec); synchronises with the device.  After crp_set_blocking(0) the counters lag behind n_exec until this
 * (or print_stat / clear_stat) is called. */
void rp_spmm_sync_stats(rp_spmm_p rp_spmm);

/* B and C may be device pointers (zero copy) or host pointers (staged through the GPU in CRP_SPMM_E2E_PANELS column panels that
 * pipeline H2D, exchange + product and D2H).  In a collective exec every rank must pass the SAME kind of buffers: the ranks cut the
 * call into the same number of exchange rounds without talking to each other. */

/* Host B / C buffers are cudaHostRegister'ed only with CRP_SPMM_PIN_HOST=1 (they must then outlive the engine);
 * this drops every registration made so far (also done by rp_spmm_free). */
void crp_unpin_host_all(void);

/* 1 if the engine was created in plan-only mode (CRP_SPMM_PLAN_ONLY=1, no device state). */
int rp_spmm_is_plan_only(rp_spmm_p rp_spmm);

/* Library version string. */
const char *crp_version(void);

#ifdef __cplusplus
}
#endif

#endif
