/*
 * oracle/tm_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A plain-C restatement of the reference's Dornic lattice time step
 * (HAL9K000/Trophic_Maelstorm, simulations/{Percolation_FD,Rietkerk_FD,Rietkerk_FPE}).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.  The product path
 * (trophic_maelstorm_b200/, include/tm_lattice.h) never links or calls it.
 *
 * Parity status: PINNED against the reference itself.  oracle/Makefile compiles the
 * reference's own sources from /root/reference into oracle/_ref/ (git-ignored) with a
 * replay URBG substituted for std::mt19937_64 (oracle/ref_shim/replay_urbg.h); the
 * SEQ stream mode below consumes the very same uniforms in the very same order, and
 * tests/golden/ holds reference outputs generated that way (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows.  Shorthands:
 *   P/ = simulations/Percolation_FD/stochasticSPDP.cpp
 *   R/ = simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp
 *   F/ = simulations/Rietkerk_FPE/rietkerk_bjork_FKE.cpp
 *   LS = libstdc++ 13 <bits/random.tcc> (the reference's sampler dependency,
 *        /usr/include/c++/13/bits/random.tcc; not vendored by the reference)
 */
#ifndef TM_ORACLE_H
#define TM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- models ------------------------------------------------------------------- */
enum {
    TMO_MODEL_DP        = 0, /* P/:940-1072   Dornic + closed form rho/(1+b rho dt)      */
    TMO_MODEL_DP_CUBIC  = 1, /* P/:1809-2062  Dornic + pointwise RK4 of -b rho^2-c rho^3 */
    TMO_MODEL_RIETKERK  = 2, /* R/:3648-4183  veg Dornic + 3-field RK4 (f_2Dor) + clamp  */
    TMO_MODEL_RIETKERK_DET = 3 /* R/:1718-1814 noise-off: full RHS f_2D, classical RK4    */
};

/* Field order everywhere: [0]=P (vegetation / rho), [1]=W (soil water), [2]=O (surface water);
 * row-major i = row*L + col, periodic (R/:1609-1628, P/:416-436). */
typedef struct tmo_params {
    int32_t model;
    int32_t L;
    double dt, dx;
    /* DP: d rho = a rho - b rho^2 (- c rho^3) + D lap rho + sigma sqrt(rho) eta */
    double a, b, c, D, sigma;
    /* Rietkerk (names as R/:3397): rainfall a is `a` above; c is `rc` */
    double rc, gmax, alpha, rW, W0, d /* = M[0] */, K1, K2;
    double D0, D1, D2;   /* D[0..2] (P, W, O) */
    double sigma0;       /* sigma[0] */
    /* clamp (R/:4176: eps=1e-12, lo=-1e-80 ; F/:4460: eps=1e-10, lo=0) */
    double clamp_eps, clamp_lo;
} tmo_params;

/* ---- uniform streams -------------------------------------------------------------- */
enum {
    TMO_SRC_SEQ     = 0, /* one sequential array of uniforms shared by all cells, visited in
                            increasing i -- the reference's mt19937_64 order.  The explicit
                            std::normal_distribution `norm` (R/:3631) keeps its cached second
                            variate across cells, as in the reference.                     */
    TMO_SRC_REPLAY  = 1, /* per-cell slice u[cell*stride + k], k < stride (host-supplied
                            uniforms; the replay contract of the product's tm_step_replay)   */
    TMO_SRC_PHILOX  = 2  /* Philox4x32-10, key = seed, counter = (cell, trial, step, block)  */
};

typedef struct tmo_rng {
    int32_t kind;
    int32_t _pad;
    const double *u;     /* SEQ / REPLAY */
    uint64_t n_u;        /* length of u */
    uint64_t pos;        /* SEQ: next index (in/out) */
    uint64_t stride;     /* REPLAY: uniforms per cell */
    uint64_t seed;       /* PHILOX */
    uint32_t trial;      /* PHILOX */
    uint32_t step;       /* PHILOX (caller increments per step) */
    uint64_t overflow;   /* out: draws past the end of u / of a cell's slice */
    uint64_t draws;      /* out: uniforms consumed */
    /* persistent state of the reference's explicit `norm` object (SEQ mode only) */
    int32_t norm_saved_available;
    int32_t _pad2;
    double norm_saved;
} tmo_rng;

/* ---- elementary functions of the sampler spec (bit-reproducible: IEEE + fma only) -- */
double tmo_log(double x);
double tmo_exp(double x);
void tmo_sincos_turn128(double a, double *sn, double *cs); /* sin, cos of 2 pi a/128, a in [0,128] */
double tmo_lgamma_int(double k);
void   tmo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double tmo_u52_from_bits(uint64_t w);

/* ---- samplers (restating LS) driven by an array of uniforms; *pos advances -------- */
double tmo_test_normal_pair(const double *u, uint64_t n, uint64_t *pos, double *second);
double tmo_test_normal_pair_philox(uint64_t seed, uint32_t cell, uint32_t trial, uint32_t step, double *second);
long   tmo_test_poisson_philox(uint64_t seed, uint32_t cell, uint32_t trial, uint32_t step, double mean);
long   tmo_test_poisson(const double *u, uint64_t n, uint64_t *pos, double mean);
double tmo_test_gamma(const double *u, uint64_t n, uint64_t *pos, double shape);

/* ---- one time step of one lattice -------------------------------------------------
 * f: 1 field (DP) or 3 fields (Rietkerk) of L*L doubles, contiguous SoA [field][i]; updated
 * in place.  Returns 0, or -1 on bad arguments. */
int tmo_step(const tmo_params *p, tmo_rng *rng, double *f);

/* Dornic sub-step only (a2/a4): writes rho* into out[L*L] from field prev[L*L]. */
int tmo_dornic_only(const tmo_params *p, tmo_rng *rng, const double *prev, double *out);

/* RK4 sub-step alone (a6 / a12) on f[3][L*L]: full_rhs = 0 f_2Dor (R/:2389), 1 f_2D (R/:1753). */
int tmo_rk4(const tmo_params *p, double *f, int full_rhs);

/* reductions (a8): sequential left-to-right sum / L^2 and #{rho>0}; R/:1211,1245 P/:123,145 */
double tmo_mean(const double *x, int64_t n);
double tmo_occupied(const double *x, int64_t n);

/* time-bin generator (a10): P/:336-378 = R/:1472-1514. Returns count written (<= cap). */
int tmo_logarithm10_time_bins(double t_max, double dt, double *out, int cap);

/* running mean / variance over surviving replicates (a9): P/:173-208.
 * acc: [n][5] = {t, mean, var, n_surv, n_active}; rep: [n][2] = {n_active, mean}; j = replicate index. */
void tmo_var_mean_incremental_surv_runs(double *acc, const double *rep, int n, int j);

/* spreading experiments: meansq_spread_of_vector P/:156 + theborderwall P/:380, and the
 * replicate accumulation spreading_incremental_surv_runs P/:230. */
void tmo_spread_stats(const double *x, int g, int c_x, int c_y, double *sq_dist, int64_t *n_active, int *border);
void tmo_spreading_incremental_surv_runs(double *acc, const double *rep, int n);

/* ---- FKE movement: oracle/tm_oracle_fke.c ---------------------------------------
 * advdiff_FKE_MultiSp F/:3614-3695, precompute_gaussian F/:1735, set_global_FKE_params F/:195-262. */
typedef struct tmo_fke_terms {
    int32_t L;
    int32_t bound_D;              /* sig_D_ScaledBounds[s] */
    int32_t v_row_lo, v_row_hi;   /* advection window rows    [mu_x - lo, mu_x + hi] */
    int32_t v_col_lo, v_col_hi;   /* advection window columns [mu_y - lo, mu_y + hi] */
    double mu_vel_prefactor;      /* F/:249 */
    double epsilon;               /* F/ header :119, 1e-10 */
} tmo_fke_terms;
typedef struct tmo_fke_derived {
    double beta_vel, sigma_D, sigma_vD, mu_vel_prefactor;
    int32_t bound_D, bound_vD_first, bound_vD_second, pad_;
} tmo_fke_derived;
void tmo_fke_precompute_gaussian(double sigma, int Rbound, double *out);
void tmo_fke_params(double D, double v, double Dt, double dx, int CI, tmo_fke_derived *o);
int tmo_fke_move(const tmo_fke_terms *t, const double *rho, const double *gamma, const double *v_eff,
                 const double *gD, const double *gV, double *out);

/* ---- consumer-side operators of the SPB = 2 model (first part): oracle/tm_oracle_consumer.c -------------
 * calc_gamma_2Sp_NonRefugia R/:2530, f_2Dor_2Sp R/:2445, RK4_Integrate_Stochastic_2Sp R/:2472. */
typedef struct tmo_cons_params {
    int32_t L, pad_;
    double dt, a, c, gmax, alpha, rW, W0, K1, K2;
    double Dw, Do;            /* Dxd2[2], Dxd2[3] = D[s]/dx^2 of soil and surface water */
    double A01, H01, E1;      /* A[0][1], H[0][1], E[1] */
} tmo_cons_params;
typedef struct tmo_graz_params {
    int32_t L, pad_;
    double dt, dx;
    double D, v, M, sigma;    /* D[s], v[s], M[s], sigma[s] of the grazer */
    double D0floor;           /* diff_coefficient[0] = D[0]/dx^2: floor of the advection-corrected diffusion, R/:4014 */
} tmo_graz_params;
void tmo_gamma_perception(int L, const double *prey, const int32_t *offsets, int eff, double rho_avg_prey, double rho_avg_self,
                          double *gamma);
int tmo_rk4_2sp(const tmo_cons_params *p, double *f);
/* Rietkerk_FPE flavours (F/ = simulations/Rietkerk_FPE/rietkerk_bjork_FKE.cpp): calc_gamma_vel_2Sp_NonRefugia F/:2534,
 * RK4_Integrate_Stochastic_2Sp over f_DorFKE_2Sp F/:2449-2531 (grazer mortality M1 in the reaction terms) */
void tmo_gamma_perception_fpe(int L, const double *prey, const int32_t *offsets, int eff, double rho_avg_prey, double rho_avg_self,
                              double *gamma);
int tmo_rk4_fpe_2sp(const tmo_cons_params *p, double M1, double *f);
/* One time step of rietkerk_DorFPE_2D_MultiSp compiled with -DSPB=2 (F/:4094-4462) on f[4][L*L] = P, G, W, O: means and cover of the
 * entry state -> perception field (F/:4117) -> vegetation Dornic (F/:4154-4211) -> four-field RK4 without movement terms (F/:4240)
 * -> FKE movement of the grazer (advdiff_FKE_MultiSp, F/:4308) -> clamp `v < eps && v >= 0 -> 0` (F/:4458).  The grazer has NO Dornic
 * update here.  gamma: state (never reset); veff [L*L][2]: the advection vectors the movement step reads (the 2-species runner
 * never assigns them: zeros); gD, gV: the two Gaussian stencils of ft; means (may be NULL): {<P>, <G>} of the entry state. */
int tmo_step_fpe_2sp(const tmo_params *pveg, const tmo_cons_params *pc, double M1, const tmo_fke_terms *ft, const double *gD,
                     const double *gV, const int32_t *central, int n_central, double fr_graz, tmo_rng *rng, double *f, double *gamma,
                     const double *veff, const double *means);
/* the noise-off DP path: one RK4 step of percNoDiff_2D (P/:549-662, fP_2D P/:516-529) in place on rho[L*L] */
int tmo_dp_noise_off_step(int L, double a, double b, double D, double dt, double dx, double *rho);
/* the grazer's Dornic update with upwind advection, R/:3908-4114 (defined in tm_oracle.c next to the samplers) */
int tmo_grazer_dornic(const tmo_graz_params *p, tmo_rng *rng, const double *prev, double *gamma, double *out, double *veff,
                      int trig_libm);
/* one full time step of the SPB = 2 runner (R/:3752-4183) on f[4][L*L] = P, G, W, O; see tm_oracle.c */
int tmo_step_2sp(const tmo_params *pveg, const tmo_graz_params *pg, const tmo_cons_params *pc, const int32_t *central, int n_central,
                 double fr_graz, tmo_rng *rng, double *f, double *gamma, double *veff, int trig_libm);
/* ... with {<P>, <G>} of the entry state supplied by the caller (NULL: the left-to-right sums) */
int tmo_step_2sp_means(const tmo_params *pveg, const tmo_graz_params *pg, const tmo_cons_params *pc, const int32_t *central, int n_central,
                       double fr_graz, tmo_rng *rng, double *f, double *gamma, double *veff, int trig_libm, const double *means);


/* ---- the SPB = 3 model (vegetation + grazer + predator): fields [0]=P, [1]=G, [2]=Pr, [3]=W, [4]=O -----------------
 * f_2Dor_3Sp R/:2607-2650, RK4_Integrate_Stochastic_MultiSp R/:2652-2720, calc_gamma_3Sp_NonRefugia R/:3153-3268, and the time-loop
 * body of rietkerk_Dornic_2D_MultiSp compiled with -DSPB=3 (R/:3752-4183: the consumers' Dornic loop R/:3908-4114 runs for s = 1, 2). */
typedef struct tmo_cons3_params {
    int32_t L, pad_;
    double dt, a, c, gmax, alpha, rW, W0, K1, K2;
    double Dw, Do;                 /* Dxd2[3], Dxd2[4] = D[s]/dx^2 of soil and surface water */
    double A01, H01, E1;           /* A[0][1], H[0][1], E[1]: grazer on vegetation */
    double A12, H12, E2;           /* A[1][2], H[1][2], E[2]: predator on grazer */
} tmo_cons3_params;
/* veg, graz, pred: DRho[0..2]; offsets: the distance-sorted disc list; new_length: its prefix after the vegetation-cover cut (R/:3162);
 * fr1, fr2: the two radius fractions; avg[3]: spatial means; gamma1, gamma2 [L*L]: OUT (assigned, not accumulated: R/:3239, 3248) */
void tmo_gamma_3sp(int L, const double *veg, const double *graz, const double *pred, const int32_t *offsets, int new_length,
                   double fr1, double fr2, const double *avg, double *gamma1, double *gamma2);
int tmo_rk4_3sp(const tmo_cons3_params *p, double *f);
/* one time step on f[5][L*L]; gamma [2][L*L] (grazer, predator); veff (may be NULL) [2][L*L][2]; means (may be NULL): {<P>, <G>, <Pr>} */
int tmo_step_3sp(const tmo_params *pveg, const tmo_graz_params *pg1, const tmo_graz_params *pg2, const tmo_cons3_params *pc,
                 const int32_t *central, int n_central, double fr1, double fr2, tmo_rng *rng, double *f, double *gamma, double *veff,
                 int trig_libm, const double *means);
int tmo_rk4_fpe_3sp(const tmo_cons3_params *p, double M1, double M2, double *f);
/* One time step of rietkerk_DorFPE_2D_MultiSp compiled with -DSPB=3 (F/:4094-4462) on f[5][L*L]: means and cover of the entry state ->
 * calc_gamma_vel_3Sp_NonRefugia (F/:3371-3483: per site and species, in the order of decreasing perception radius, ONE uniform for a new
 * advection direction -- the vegetation's draw included -- then the perception field, ACCUMULATED onto its old value, a vanished
 * consumer getting gamma = 1) -> vegetation Dornic -> five-field RK4 over f_DorFKE_3Sp -> FKE movement of grazer and predator -> clamp.
 * M, v: {grazer, predator}; ft[2], gD[2], gV[2]: their movement terms and stencils; gamma [2][L*L] and veff [2][L*L][2] are state.
 * Per-cell streams: the direction uniforms of a cell come from their own stream (PHILOX: seed xor 0xA0761D6478BD642F, uniform k for the
 * k-th species in radius order; REPLAY: the slice after the vegetation's). */
int tmo_step_fpe_3sp(const tmo_params *pveg, const tmo_cons3_params *pc, const double *M, const double *v, const tmo_fke_terms *ft,
                     const double *const *gD, const double *const *gV, const int32_t *central, int n_central, double fr1, double fr2,
                     tmo_rng *rng, double *f, double *gamma, double *veff, int trig_libm, const double *means);

#ifdef __cplusplus
}
#endif
#endif
