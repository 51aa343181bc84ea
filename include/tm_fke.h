/*
 * include/tm_fke.h -- C ABI of the FKE movement step (analytic Fokker-Planck advection-diffusion of the
 * consumer species) on B200 (sm_100a).  Plain pointers and sizes only.
 *
 * What it replaces (F/ = simulations/Rietkerk_FPE/rietkerk_bjork_FKE.cpp):
 *   launchAdvDiffKernel_MultiSp(double* Rho_t, double* Rho_tsar, double* gamma, std::pair<double,double>* v_eff,
 *                               double* gaussian_stencil_D, double* gaussian_stencil_VXY, int L)
 *                                        simulations/Rietkerk_FPE/rietkerk_bjork_basic.h:457-459,
 *                                        kernel_advdiff_FKE2D.cu:92-180 (updateMovement), :222-235
 *   copyToDeviceConstantMemory_AdvTerms(const int* sig_D_ScaledBounds, const int2* sigma_vD_ScBounds,
 *                               const double* mu_vel_prefactor, const int* size_gauss_D, const int* size_gauss_VXY)
 *                                        rietkerk_bjork_basic.h:466, kernel_advdiff_FKE2D.cu:28-88
 *   advdiff_FKE_MultiSp(...)             F/:3614-3695 (the CPU path the GPU kernel mirrors)
 * and the per-step cudaMemcpy / cudaMemset choreography around them, F/:4282-4299.
 *
 * Differences by design: the reference keeps its window bounds in process-global __constant__ memory (one set for the
 * whole process); here they belong to a tm_fke_ctx (one per calling thread).  The reference scatters with FP64 atomics
 * (sum order varies run to run); this library gathers in a fixed order, so results are bit-reproducible and agree with
 * the reference's CPU path to a few ulp.  No CPU fallback: without a CUDA device tm_fke_create fails with TM_E_CUDA.
 */
#ifndef TM_FKE_H
#define TM_FKE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Movement terms of ONE species (the reference's arrays indexed by s >= 1). */
typedef struct tm_fke_terms {
    int32_t bound_D;              /* sig_D_ScaledBounds[s]: diffusion window c +- bound_D          F/:244  */
    int32_t v_row_lo, v_row_hi;   /* advection window rows    [mu_x - lo, mu_x + hi]                        */
    int32_t v_col_lo, v_col_hi;   /* advection window columns [mu_y - lo, mu_y + hi]                        */
                                  /*   CPU path F/:3640-3641: lo = sig_vD_ScaledBounds[s].first, hi = .second on both axes */
                                  /*   CUDA path .cu:126-127: rows +- .x, columns +- .y                     */
    int32_t n_gauss_D;            /* entries of gaussian_stencil_D[s]   (>= bound_D + 1)                    */
    int32_t n_gauss_V;            /* entries of gaussian_stencil_VXY[s] (>= the largest of the four v_* + 1) */
    int32_t pad_;
    double  mu_vel_prefactor;     /* F/:249 */
    double  epsilon;              /* sites with rho < epsilon do not move (and vanish)   F/:3622, header :119 */
} tm_fke_terms;

typedef struct tm_fke_ctx tm_fke_ctx;

/* L x L periodic lattice, n_species moving species (grazer[, predator]); binds `device` and one stream.
 * The stream is a blocking one: it orders itself against the legacy default stream like the reference's launch
 * (cudaMemset before it, cudaMemcpy after it, F/:4292-4299). */
int tm_fke_create(int L, int n_species, int device, tm_fke_ctx **out);
int tm_fke_destroy(tm_fke_ctx *);
const char *tm_fke_last_error(const tm_fke_ctx *); /* ctx == NULL: error of the last failed tm_fke_create on this thread */

/* Window bounds, prefactor and the two Gaussian stencils (host pointers) of species s in [0, n_species). */
int tm_fke_set_terms(tm_fke_ctx *, int s, const tm_fke_terms *t, const double *gauss_D, const double *gauss_V);

/* The same with the stencils already resident on the device (the reference uploads them once per run, F/:3775-3800). */
int tm_fke_set_terms_device(tm_fke_ctx *, int s, const tm_fke_terms *t, const double *d_gauss_D, const double *d_gauss_V);

/* One movement step of every species with HOST buffers: rho, gamma, rho_out [n_species][L*L]; v_eff [n_species][L*L][2]
 * ({first, second} = {row, column} velocity, the memory layout of std::pair<double,double>).  Uploads, moves, downloads. */
int tm_fke_move_host(tm_fke_ctx *, const double *rho, const double *gamma, const double *v_eff, double *rho_out);

/* The same with DEVICE pointers (the launchAdvDiffKernel_MultiSp seam: the caller owns the buffers); asynchronous on the
 * context's stream, d_rho_out need not be zeroed. */
int tm_fke_move_device(tm_fke_ctx *, const double *d_rho, const double *d_gamma, const double *d_v_eff, double *d_rho_out);

int tm_fke_sync(tm_fke_ctx *);
void *tm_fke_cuda_stream(tm_fke_ctx *);
/* Device time (ms, CUDA events on the ctx stream) of the kernels of the last move call, and how many launches that was. */
int tm_fke_last_ms(tm_fke_ctx *, float *ms, int *n_launches);
/* max |displacement| (rows, columns) found by the last move of species s: the gather radius it used.  A displacement beyond 64 cells
 * on either axis (|v_eff * mu_vel_prefactor| > 64: three orders of magnitude above what the reference's constants give) is refused by
 * the column pass: the moved field of that species comes back as NaN and this call reports the offending radius.  The reference's
 * scatter (kernel_advdiff_FKE2D.cu:119-121) has no such limit; the gather form would visit (2 R + 1)^2 sources per target. */
int tm_fke_last_displacement(tm_fke_ctx *, int s, int32_t out[2]);

#ifdef __cplusplus
}
#endif
#endif
