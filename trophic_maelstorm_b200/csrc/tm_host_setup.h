// tm_host_setup.h -- host-side derivation of kernel constants and launch geometry from tm_params.
// Shared by the C-ABI layer (tm_lattice.cu) and the host-thread emulator used by the CPU tests
// (tests/emu), so both drive the kernels with identical arguments.
#pragma once
#include "../../include/tm_lattice.h"
#include "tm_kernels.cuh"
#include "tm_rk4.cuh"
#include <cmath>

namespace tml {

static inline bool is_riet(int m) { return m == TM_MODEL_RIETKERK || m == TM_MODEL_RIETKERK_DET; }

// Dornic constants exactly as the reference computes them on the host:
// P/:952-954 (DP), R/:3461-3463 (Rietkerk vegetation).
static inline DornicConsts dornic_consts(const tm_params &p, double a)
{
    DornicConsts k{};
    double beta, lambda_exp;
    if (is_riet(p.model)) {
        const double dx2 = p.dx * p.dx;
        beta = -p.d - 4 * p.D0 / dx2;
        lambda_exp = exp(beta * p.dt);
        k.lambda = 2 * beta / (p.sigma0 * p.sigma0 * (lambda_exp - 1.0));
        k.diffc = p.D0 / dx2;
        k.sigma2 = p.sigma0 * p.sigma0;
        k.sigma2_1 = 1 / (p.sigma0 * p.sigma0);
    } else {
        k.diffc = p.D / (p.dx * p.dx);
        beta = a - 4 * p.D / (p.dx * p.dx);
        lambda_exp = exp(beta * p.dt);
        k.lambda = 2 * beta / (p.sigma * p.sigma * (lambda_exp - 1.0));
        k.sigma2 = p.sigma * p.sigma;
        k.sigma2_1 = 1 / k.sigma2;
    }
    k.ll = k.lambda * lambda_exp;
    return k;
}

static inline RietConsts riet_consts(const tm_params &p)
{
    RietConsts c{};
    const double dx2 = p.dx * p.dx;
    c.dt = p.dt; c.dt2 = p.dt / 2.0; c.dt6 = p.dt / 6.0; // R/:170
    c.cgmax = p.rc * p.gmax;                             // R/:2365
    c.gmax = p.gmax; c.alpha = p.alpha; c.K1 = p.K1; c.K2 = p.K2; c.K2W0 = p.K2 * p.W0; c.rW = p.rW;
    c.Dp = p.D0 / dx2; c.Dw = p.D1 / dx2; c.Do = p.D2 / dx2; // R/:3463-3480
    c.d = p.d; c.rc = p.rc; c.W0 = p.W0;
    c.eps = p.clamp_eps; c.lo = p.clamp_lo;
    c.dk = dornic_consts(p, p.a);
    c.inv_lambda = 1.0 / c.dk.lambda;
    c.mu_scale = 2.0 * c.dk.diffc * c.dk.sigma2_1;
    c.dt2cg = c.dt2 * c.cgmax; c.dtcg = c.dt * c.cgmax; c.dt6cg = c.dt6 * c.cgmax;
    c.aKap = c.alpha * c.K2 * (1.0 - c.W0);
    c.aK2W0 = c.alpha * c.K2W0; c.cP = -(c.d + 4 * c.Dp); c.cW = -(c.rW + 4 * c.Dw); c.cO = -4 * c.Do;
    return c;
}

// strip width: one thread per column incl. 4+4 halo columns; maximise L / (ns * nthreads)
static inline void choose_strips(int L, int *ns_out, int *tx_out, int *nthreads_out)
{
    double best = -1; int bns = 1, btx = L, bnt = 32;
    for (int nt = 64; nt <= kMaxThreads; nt += 32) {
        int tx = (nt - 2 * kHalo) & ~1;
        if (tx > L) tx = (L + 1) & ~1;
        const int ns = (L + tx - 1) / tx;
        tx = (((L + ns - 1) / ns) + 1) & ~1;             // even out the strips
        const int need = ((tx + 2 * kHalo + 31) / 32) * 32;
        const double eff = (double)L / ((double)ns * need);
        if (eff > best + 1e-12) { best = eff; bns = ns; btx = tx; bnt = need; }
    }
    *ns_out = bns; *tx_out = btx; *nthreads_out = bnt;
}


// the ten Philox4x32-10 round keys of a seed
static inline PhiloxKeys philox_keys(unsigned long long seed)
{
    PhiloxKeys rk;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) { rk.k[r][0] = k0; rk.k[r][1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return rk;
}

// FAST-mode log table: c_i = 1 + (i + 1/2)/128, {1/c_i, log c_i}
static inline void make_log_table(double *tab /* [128][2] */)
{
    for (int i = 0; i < 128; i++) { const double ci = 1.0 + (i + 0.5) / 128.0; tab[2 * i] = 1.0 / ci; tab[2 * i + 1] = std::log(ci); }
}

// fills the layout part of Geom for a slab [row0,row0+nrows) of an L x L lattice
static inline void make_geom(const tm_params &p, int n_trials, int row0, int nrows, Geom *g)
{
    g->L = p.L; g->nrows = nrows; g->row0 = row0; g->wrap = (nrows == p.L) ? 1 : 0;
    g->rows_alloc = g->wrap ? nrows : nrows + 2 * kHalo;
    g->n_trials = n_trials; g->n_fields = is_riet(p.model) ? 3 : 1;
    choose_strips(g->L, &g->ns, &g->TX, &g->nthreads);
    int need_cols = g->ns * g->TX + 2 * kHalo + 2;
    {   // the two-column RK4 kernel (tm_rk4.cuh) has its own strips: its last strip must end inside the padded row
        int ns2, tx2;
        choose_strips2(g->L, &ns2, &tx2);
        if (ns2 * tx2 + 2 * kHalo > need_cols) need_cols = ns2 * tx2 + 2 * kHalo;
    }
    g->Lp = ((need_cols > g->L + 2 * kHalo ? need_cols : g->L + 2 * kHalo) + 15) / 16 * 16;
    g->field_stride = (long long)g->rows_alloc * g->Lp;
    g->trial_stride = g->field_stride * g->n_fields;
}

// arguments and grid of the two-column RK4 kernel (n_cta_slots = SMs x resident CTAs per SM)
static inline void rk4_launch_shape(const Geom &g, int n_cta_slots, Rk4Args *A, int *grid)
{
    choose_strips2(g.L, &A->ns2, &A->TX2);
    const long long total = (long long)g.n_trials * A->ns2 * g.nrows;
    long long gr = total / 16;                       // at least 16 rows per CTA: 8 of every segment's rows are recomputed halo
    if (gr < 1) gr = 1;
    if (gr > n_cta_slots) gr = n_cta_slots;
    *grid = (int)gr;
}

// grid of the persistent Rietkerk kernel (n_cta_slots = SMs x resident CTAs per SM)
static inline void riet_launch_shape(const Geom &g, int n_cta_slots, int *grid)
{
    const long long total = (long long)g.n_trials * g.ns * g.nrows;
    long long gr = total / 16;
    if (gr < 1) gr = 1;
    if (gr > n_cta_slots) gr = n_cta_slots;
    *grid = (int)gr;
}

} // namespace tml
