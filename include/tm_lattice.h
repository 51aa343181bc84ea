/*
 * include/tm_lattice.h -- the drop-in boundary: a C ABI over the B200 (sm_100a) Dornic lattice
 * integrator.  Plain pointers and sizes only; no STL, no torch types.
 *
 * What it replaces.  The reference (HAL9K000/Trophic_Maelstorm) has no C ABI; its seams for this
 * path are the C++ replicate runners and the per-step numerics they inline:
 *
 *   P/ = simulations/Percolation_FD/stochasticSPDP.cpp      R/ = simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp
 *   F/ = simulations/Rietkerk_FPE/rietkerk_bjork_FKE.cpp
 *
 *   percolationDornic_2D            P/:940   (SPDP.h:96)     -> TM_MODEL_DP
 *   expanded_percolationDornic_2D   P/:1809  (SPDP.h:117)    -> TM_MODEL_DP_CUBIC
 *   rietkerk_Dornic_2D_MultiSp      R/:3397  (rietkerk_bjork_basic.h:388), SPB=1 -> TM_MODEL_RIETKERK
 *   rietkerk_DorFPE_2D_MultiSp      F/:3697, SPB=1 (same step, clamp 1e-10/0)    -> TM_MODEL_RIETKERK
 *   RK4_Wrapper_2D / RK4_Integrate  R/:1816 / R/:1753 (noise-off)                -> TM_MODEL_RIETKERK_DET
 *
 * The existing GPU seam this mirrors in shape (raw device state owned behind the call, host
 * std::vectors stay the source of truth) is launchAdvDiffKernel_MultiSp,
 * simulations/Rietkerk_FPE/rietkerk_bjork_basic.h:457-467.
 *
 * Threading: one tm_ctx per calling thread (the reference calls its runners from inside
 * `#pragma omp parallel for`, R/:4509); a ctx binds one device and one stream and holds no
 * process-global mutable state.  Every call returns 0 on success or a negative TM_E_* code;
 * tm_last_error() gives the message.  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with TM_E_CUDA.
 */
#ifndef TM_LATTICE_H
#define TM_LATTICE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TM_ABI_VERSION 1

enum {
    TM_MODEL_DP           = 0, /* Dornic + rho/(1+b rho dt)                      P/:1032-1068 */
    TM_MODEL_DP_CUBIC     = 1, /* Dornic + pointwise RK4 of -b rho^2 - c rho^3   P/:1925-2014, :1775 */
    TM_MODEL_RIETKERK     = 2, /* veg Dornic + RK4(f_2Dor) + clamp               R/:3752-3809, :2389, :4150 */
    TM_MODEL_RIETKERK_DET = 3  /* RK4(f_2D), no sampling, no clamp               R/:1718-1814 */
};

enum {
    TM_OK = 0,
    TM_E_ARG = -1,     /* bad argument */
    TM_E_CUDA = -2,    /* CUDA runtime/driver error (includes "no device") */
    TM_E_STATE = -3,   /* call sequence error (e.g. step before upload) */
    TM_E_NUMERIC = -4  /* non-finite value produced (reference: dump frames and exit(3), R/:4206) */
};

/* Physical parameters.  Field order everywhere: [0]=P (vegetation, or the single DP field rho),
 * [1]=W (soil water), [2]=O (surface water) -- the reference's SPB=1 memory order (R/:3517). */
typedef struct tm_params {
    int32_t model;
    int32_t L;             /* lattice side `g`; L >= 8 */
    double dt, dx;
    /* DP:  d rho = a rho - b rho^2 (- c rho^3) + D lap rho + sigma sqrt(rho) eta   (P/:940) */
    double a;              /* also the Rietkerk rainfall `a` (R/:3397) */
    double b, c, D, sigma;
    /* Rietkerk (argument names of R/:3397; `rc` is its `c`, `d` is M[0]) */
    double rc, gmax, alpha, rW, W0, d, K1, K2;
    double D0, D1, D2;     /* D[0..2] */
    double sigma0;         /* sigma[0] */
    double clamp_eps;      /* R/:3402 1e-12   F/: 1e-10 */
    double clamp_lo;       /* R/:4176 -1e-80  F/:4460 0 */
} tm_params;

typedef struct tm_ctx tm_ctx;

/* Per-trial reductions of the current state (a8: mean_of_vector / occupied_sites_of_vector,
 * R/:1211,1245), plus the guards the reference evaluates every step (R/:4156-4174, P/:2022-2041). */
typedef struct tm_stats {
    double   mean[3];       /* <field>_x = sum / L^2 */
    uint64_t active[3];     /* #{field > 0} */
    uint64_t nonfinite;     /* NaN/Inf cells in any field */
    uint64_t negative;      /* cells < 0 in any field */
} tm_stats;

/* --- lifecycle ------------------------------------------------------------------------ */
int tm_abi_version(void);
int tm_device_count(void);                  /* 0 when no CUDA device is visible */
/* Sizing aids of the ensemble scheduler (the reference's `#pragma omp parallel for` over a-values,
 * Rietkerk_FD/rietkerk_bjork_basic.cpp:4509-4542, has no memory limit to respect; a device context does):
 * free / total memory of a device, and the device bytes a context of n_trials whole lattices of *p will allocate. */
int tm_device_mem(int device, uint64_t *free_bytes, uint64_t *total_bytes);
int tm_ctx_bytes(const tm_params *p, int n_trials, uint64_t *bytes);
const char *tm_last_error(const tm_ctx *);  /* ctx == NULL: error of the last failed tm_create on this thread */

/* n_trials independent lattices (replicates and/or a-values) advanced by every launch.
 * row0/nrows select a row slab [row0,row0+nrows) of the L x L lattice for domain decomposition;
 * pass 0 and L for a whole lattice. */
int tm_create(const tm_params *p, int n_trials, int device, tm_ctx **out);
int tm_create_slab(const tm_params *p, int n_trials, int device, int row0, int nrows, tm_ctx **out);
int tm_destroy(tm_ctx *);

/* Per-trial control parameter `a` (the a-sweep of first_order_critical_exp_delta*, P/:2470, R/:4519);
 * defaults to p->a for every trial. */
int tm_set_trial_a(tm_ctx *, int trial, double a);

/* strict != 0: keep the reference's floating-point operation order in the RK4 stages (bit-identical
 * to the oracle); 0 (default): the fast arrangement (shared reciprocal, fused multiply-adds). */
int tm_set_strict(tm_ctx *, int strict);

/* Rietkerk model only.  fused == 0 (default): each step is two launches -- the vegetation Dornic sweep (one
 * thread per column, 32 warps per SM) into a rho* scratch field, then the marching RK4 kernel; 64 B of HBM
 * traffic per site.  fused != 0: one launch per step, 48 B per site, slower on B200 (both kernels are bound
 * by FP64/issue slots, not HBM; the fused loop body does not fit the instruction cache once unrolled).
 * Results are bit-identical either way. */
int tm_set_fused(tm_ctx *, int fused);

/* --- state ----------------------------------------------------------------------------- */
/* host_soa: double[n_trials][n_fields][nrows][L], n_fields = 1 (DP) or 3 (Rietkerk).
 * Pinned host memory makes these asynchronous-capable; pageable memory works too. */
int tm_upload(tm_ctx *, const double *host_soa);
int tm_download(tm_ctx *, double *host_soa);
int tm_n_fields(const tm_ctx *);

/* --- stepping -------------------------------------------------------------------------- */
/* Philox4x32-10 stream: key = seed, counter = (global cell index, trial0 + trial, step, block). */
int tm_seed(tm_ctx *, uint64_t seed, uint32_t trial0);
int tm_set_step_counter(tm_ctx *, uint64_t step);
uint64_t tm_step_counter(const tm_ctx *);

/* nsteps full time steps; statistics of the final state are accumulated in the last launch. */
int tm_step(tm_ctx *, int nsteps);

/* Replay: ONE step in which cell i of trial t draws its k-th uniform from
 * host_uniforms[(t*L*L + i)*stride + k] instead of Philox.  Draws past `stride` are counted in
 * tm_sampler_overflow() and return 0.75. */
int tm_step_replay(tm_ctx *, const double *host_uniforms, uint64_t stride);

/* The Dornic sub-step alone (a2/a4): rho* into host_out[n_trials][nrows][L]; state unchanged. */
int tm_dornic_only(tm_ctx *, double *host_out);

/* Slab halos: pack/unpack the 4 boundary rows x n_fields of every trial (device pointers,
 * for NCCL or peer copies by the caller); `which` 0 = top rows, 1 = bottom rows. */
int tm_halo_rows(void);
int tm_halo_bytes(const tm_ctx *, uint64_t *bytes);
int tm_halo_pack(tm_ctx *, int which, void **dev_buf);
int tm_halo_unpack(tm_ctx *, int which, const void *dev_buf);
void *tm_cuda_stream(tm_ctx *);

/* Overlapped slab stepping (BASELINE config 5: "Rietkerk L=16384 single trial domain-decomposed across 2/4/8 GPUs with
 * NVLink halo exchange"; the reference has no counterpart, its lattice lives in one address space).  One step =
 *   tm_slab_step_begin : the two boundary bands (tm_slab_band_rows() rows each) of the NEW state, on tm_slab_comm_stream()
 *   <transport>        : on that stream, send the new edge rows / receive the new ghost rows (tm_slab_halo_ptrs, next = 1);
 *                        include/tm_slab.h does this with ncclSend / ncclRecv
 *   tm_slab_step_end   : the inner rows on the main stream, concurrently with the two items above.
 * tm_slab_join makes the main stream wait for the band stream before tm_measure / tm_download / tm_step.
 * Results are bit-identical to tm_step on the undivided lattice. */
int tm_slab_band_rows(void);
int tm_state_buffers(tm_ctx *, void **buf0, void **buf1, uint64_t *bytes, int *cur, int *row_doubles, int *rows_alloc);
void *tm_slab_comm_stream(tm_ctx *);
int tm_slab_reserve_sms(tm_ctx *, int n_sm);
int tm_slab_halo_ptrs(tm_ctx *, int next, int side, int trial, int field, void **send_ptr, void **recv_ptr, uint64_t *bytes);
int tm_slab_step_begin(tm_ctx *);
int tm_slab_step_end(tm_ctx *);
int tm_slab_join(tm_ctx *);
/* Copy-engine transport for slab contexts of one process (cudaMemcpyPeerAsync over NVLink): dst fetches its ghost rows from its
 * two ring neighbours.  next != 0: inside an open overlapped step of all three (between _begin and _end); 0: the current state. */
int tm_slab_pull_halos(tm_ctx *dst, tm_ctx *up, tm_ctx *down, int next);

/* --- measurement ----------------------------------------------------------------------- */
/* out[n_trials]; blocks until the stream is idle. */
int tm_measure(tm_ctx *, tm_stats *out);
/* Spreading-experiment observables of field 0 about the seed site (c_x, c_y) = (row, column): what the reference's
 * spreading_exp_Dornic_2D evaluates every step (meansq_spread_of_vector P/:156-169, theborderwall P/:380-390).
 * out[n_trials]; blocks until the stream is idle. */
typedef struct tm_spread {
    uint64_t n_active;      /* N(t) = #{rho > 0} */
    double   sq_dist;       /* sum over active sites of (p - c_x)^2 + (q - c_y)^2; R^2(t) = sq_dist / N */
    int32_t  border;        /* 1 when an active site lies on the first/last row or column */
    int32_t  pad_;
} tm_spread;
int tm_measure_spread(tm_ctx *, int c_x, int c_y, tm_spread *out);
int tm_sampler_overflow(tm_ctx *, uint64_t *count);
int tm_sync(tm_ctx *);

/* Device time of the kernels launched by the last tm_step()/tm_step_replay() call, in ms
 * (CUDA events on the ctx stream), and how many kernels that was. */
int tm_last_step_ms(tm_ctx *, float *ms, int *n_launches);

/* Rietkerk kernel pair only (measurement aid).  tm_set_kernel_timing(ctx, 1): the following tm_step() calls also
 * record a CUDA event between the two kernels of each of their first 64 steps; tm_kernel_ms() then returns the device
 * time summed over those steps of the Dornic sweep (ms[0]) and of the RK4 marching kernel (ms[1]) and the number
 * of steps they cover.  Off by default: the extra events are not part of a production step. */
int tm_set_kernel_timing(tm_ctx *, int on);
int tm_kernel_ms(tm_ctx *, double ms[2], int *n_steps);

/* --- one-call host-buffer form (what a reference runner would call per measurement window):
 * upload host_soa -> nsteps -> stats (-> download when download != 0). */
int tm_run_host(tm_ctx *, double *host_soa, int nsteps, int download, tm_stats *out);

/* The noise-off DP path: nsteps classical RK4 steps of drho/dt = a rho - b rho^2 + D lap4(rho) (fourth-order nine-point-cross
 * Laplacian, periodic) in place on rho[L*L] -- replaces the time-loop body of percNoDiff_2D, Percolation_FD/stochasticSPDP.cpp:
 * 549-662 (fP_2D :516-529, determine_neighbours_R2 :394-414).  Self-contained: uploads the host field, steps on `device`,
 * downloads; ms (may be NULL) receives the device time of the 4*nsteps launches.  Bit-identical to the oracle. */
int tm_dp_noise_off_host(int device, int L, double a, double b, double D, double dt, double dx, double *rho, int nsteps, float *ms);

#ifdef __cplusplus
}
#endif
#endif
