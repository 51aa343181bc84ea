/* tm_slab.h -- C ABI of libtm_slab.so: one huge Rietkerk lattice cut into row slabs, one slab per GPU, the per-step halo rows
 * exchanged with ncclSend / ncclRecv over NVLink and hidden behind the inner rows' update (BASELINE config 5).
 *
 * The reference has no counterpart to replace: Rietkerk_FD/rietkerk_bjork_basic.cpp keeps a lattice in one address space
 * (rietkerk_Dornic_2D_MultiSp R/:3392) and parallelises over a-values only (R/:4509).  This layer sits beside tm_step for lattices
 * one GPU should not (or cannot) hold; a slab ring of any size reproduces tm_step on the undivided lattice bit for bit,
 * because every cell's Philox counter is its GLOBAL index.
 *
 * Two ways to form the ring:
 *   - one process per GPU (torchrun, MPI ...): rank 0 calls tm_slab_unique_id, the id travels by any side channel, every rank
 *     calls tm_slab_create (collective);
 *   - one process, several GPUs (the C++ host layer): tm_slab_create_local, then tm_slab_run_local drives all slabs.
 * libtm_slab.so links libtm_lattice.so (kernels, contexts) and libnccl.so.2; libtm_lattice.so itself has no NCCL dependency.
 */
#ifndef TM_SLAB_H
#define TM_SLAB_H
#include "tm_lattice.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct tm_slab tm_slab;
#define TM_SLAB_ID_BYTES 128

/* modes of tm_slab_run */
#define TM_SLAB_OVERLAP 0   /* boundary bands first, halo traffic behind them, inner rows concurrently (production) */
#define TM_SLAB_SERIAL  1   /* exchange, then the whole slab: the round-1 schedule, kept as the comparison */
#define TM_SLAB_NOCOMM  2   /* the overlapped schedule without the transfers: timing aid (results wrong at the slab edges) */

int tm_slab_unique_id(void *id /* TM_SLAB_ID_BYTES */);
/* rows of rank r of n: contiguous, sizes differ by at most one */
void tm_slab_partition(int L, int n_ranks, int rank, int *row0, int *nrows);
int tm_slab_create(const tm_params *p, int n_trials, int device, int rank, int n_ranks, const void *id, tm_slab **out);
/* transport of a local ring: NCCL send/recv, or peer copies on the copy engines (tm_slab_pull_halos; a device may then hold
 * several slabs, which is how the one-GPU tests run a ring) */
#define TM_SLAB_NCCL 0
#define TM_SLAB_PEER_COPY 1
int tm_slab_create_local(const tm_params *p, int n_trials, int n_devices, const int *devices, int transport, tm_slab **out /* [n_devices] */);
/* Transport of a ring made by tm_slab_create (collective: every rank makes the same call).  TM_SLAB_NCCL (default): ncclSend /
 * ncclRecv.  TM_SLAB_PEER_COPY: the neighbours' buffers mapped through CUDA IPC, halo rows and arrival flags moved by the copy
 * engines, arrival awaited by a stream memory operation: no kernel, so nothing competes with the inner launch for CTA slots. */
int tm_slab_set_transport(tm_slab *, int transport);
int tm_slab_destroy(tm_slab *);
const char *tm_slab_last_error(const tm_slab *); /* NULL: error of the last failed create on this thread */

tm_ctx *tm_slab_ctx(tm_slab *);                  /* the slab's context: tm_upload / tm_download / tm_measure / tm_seed / tm_set_trial_a */
int tm_slab_rows(const tm_slab *, int *row0, int *nrows);

/* nsteps Dornic + RK4 steps of the ring.  tm_slab_run: this process holds one slab (collective over the ranks);
 * tm_slab_run_local: this process holds all n slabs.  Asynchronous; tm_slab_sync waits. */
int tm_slab_run(tm_slab *, int nsteps, int mode);
int tm_slab_run_local(tm_slab **, int n, int nsteps, int mode);
int tm_slab_sync(tm_slab *);
/* device time of the last run on this slab's GPU (CUDA events on its main stream around the run, the join included);
 * bytes this rank sends per step (= receives) */
int tm_slab_last_ms(tm_slab *, float *ms);
int tm_slab_comm_bytes_per_step(const tm_slab *, uint64_t *bytes);

#ifdef __cplusplus
}
#endif
#endif
