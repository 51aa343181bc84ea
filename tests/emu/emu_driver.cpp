// tests/emu/emu_driver.cpp -- TEST TOOL.  C entry points that push one lattice through the
// product's kernel source on host threads (see tm_cpu_emu.h), with the same argument set-up the
// C-ABI layer uses (csrc/tm_host_setup.h).  Used by tests/test_kernel_emulation.py to check the
// marching pipeline, halo logic and decomposition invariance against the oracle without a GPU.
#include "tm_host_setup.h"
#include "tm_fke.cuh"
#include "tm_consumer.cuh"
#include "tm_consumer3.cuh"
#include "tm_dpdet.cuh"

#include <cmath>
#include <vector>

using namespace tml;

namespace {
void fill_ghosts(const Geom &g, std::vector<double> &buf)
{
    GhostArgs A{g, buf.data()};
    emu_launch(ghost_cols_kernel, A, 1, 64, 0);
}
void pack(const Geom &g, const double *host, std::vector<double> &buf)
{
    const int roff = g.wrap ? 0 : kHalo;
    for (int t = 0; t < g.n_trials; t++)
        for (int f = 0; f < g.n_fields; f++)
            for (int r = 0; r < g.nrows; r++)
                for (int q = 0; q < g.L; q++)
                    buf[(size_t)t * g.trial_stride + (size_t)f * g.field_stride + (size_t)(r + roff) * g.Lp + kHalo + q] =
                        host[(((size_t)t * g.n_fields + f) * g.nrows + r) * g.L + q];
}
void unpack(const Geom &g, const std::vector<double> &buf, double *host)
{
    const int roff = g.wrap ? 0 : kHalo;
    for (int t = 0; t < g.n_trials; t++)
        for (int f = 0; f < g.n_fields; f++)
            for (int r = 0; r < g.nrows; r++)
                for (int q = 0; q < g.L; q++)
                    host[(((size_t)t * g.n_fields + f) * g.nrows + r) * g.L + q] =
                        buf[(size_t)t * g.trial_stride + (size_t)f * g.field_stride + (size_t)(r + roff) * g.Lp + kHalo + q];
}
}

extern "C" {

static int g_split = 0;
void emu_set_split(int s) { g_split = s; } // 1: Dornic kernel + kSrcPre marching kernel; 2: Dornic kernel + riet_rk4_kernel (the production pair)

// fields: [n_trials][n_fields][L][L] in/out (whole lattice).  grid <= 0: as many CTAs as an n_sm-SM
// device would get.  stats_out: [n_trials][8] = sum-derived mean[3], active[3], nonfinite, negative.
int emu_step(const tm_params *p, int n_trials, double *fields, int nsteps, int strict, int grid_override,
             int n_sm, unsigned long long seed, unsigned trial0, unsigned step0,
             const double *replay_u, unsigned stride, double *stats_out, unsigned long long *counters_out)
{
    Geom g{};
    make_geom(*p, n_trials, 0, p->L, &g);
    std::vector<double> bufs[2];
    bufs[0].assign((size_t)g.n_trials * g.trial_stride, 0.0);
    bufs[1].assign((size_t)g.n_trials * g.trial_stride, 0.0);
    pack(g, fields, bufs[0]);
    fill_ghosts(g, bufs[0]);
    unsigned long long counters[4] = {0, 0, 0, 0};
    std::vector<double> a_trial(n_trials, p->a);
    int cur = 0;
    std::vector<PartialStats> partials;
    int n_records = 0;
    for (int s = 0; s < nsteps; s++) {
        const bool stats = (s == nsteps - 1);
        if (is_riet(p->model)) {
            int grid;
            riet_launch_shape(g, n_sm, &grid);
            if (grid_override > 0) grid = grid_override;
            StepArgs A{};
            A.g = g; A.c = riet_consts(*p);
            A.in = bufs[cur].data(); A.out = bufs[cur ^ 1].data();
            A.rk = philox_keys(seed);
            static double logtab[256]; make_log_table(logtab); A.logtab = logtab;
            A.a_trial = a_trial.data(); A.seed = seed; A.trial0 = trial0; A.step = step0 + s;
            A.replay_u = replay_u; A.stride = stride;
            A.counters = counters; A.dornic_out = nullptr;
            const bool det = p->model == TM_MODEL_RIETKERK_DET;
            const size_t smem = riet_smem_bytes(g.nthreads, det);
            if (det) {
                if (strict) emu_launch(riet_step_kernel<true, true, kSrcPhilox, false>, A, grid, g.nthreads, smem);
                else        emu_launch(riet_step_kernel<true, false, kSrcPhilox, false>, A, grid, g.nthreads, smem);
            } else if (replay_u) {
                if (strict) emu_launch(riet_step_kernel<false, true, kSrcReplay, false>, A, grid, g.nthreads, smem);
                else        emu_launch(riet_step_kernel<false, false, kSrcReplay, false>, A, grid, g.nthreads, smem);
            } else if (g_split) {
                std::vector<double> pstar((size_t)g.n_trials * g.field_stride, 0.0);
                DornicArgs D{};
                D.g = g; D.dk = A.c.dk; D.inv_lambda = A.c.inv_lambda; D.mu_scale = A.c.mu_scale; D.in = A.in; D.pstar = pstar.data();
                D.seed = seed; D.trial0 = trial0; D.step = step0 + s; D.rk = A.rk; D.logtab = A.logtab; D.row_lo = 0; D.row_hi = g.nrows; D.rows_per_cta = 7;
                const size_t dsmem = riet_dornic_smem_bytes();
                if (g_split == 2) { // the production pair uses the two-column sweep
                    if (strict) emu_launch(riet_dornic2_kernel<true>, D, 3, kDornic2Threads, dsmem);
                    else        emu_launch(riet_dornic2_kernel<false>, D, 3, kDornic2Threads, dsmem);
                } else
                if (strict) emu_launch(riet_dornic_kernel<true>, D, 3, kDornicThreads, dsmem);
                else        emu_launch(riet_dornic_kernel<false>, D, 3, kDornicThreads, dsmem);
                A.pstar = pstar.data();
                if (g_split == 2) { // the production pair: Dornic sweep + the two-column RK4 kernel (tm_rk4.cuh)
                    Rk4Args R{};
                    R.g = g; R.c = A.c; R.in = A.in; R.pstar = pstar.data(); R.out = A.out; R.a_trial = A.a_trial; R.counters = counters;
                    int grid2;
                    rk4_launch_shape(g, n_sm, &R, &grid2);
                    if (grid_override > 0) grid2 = grid_override;
                    if (strict) emu_launch(riet_rk4_kernel<true>, R, grid2, kRkThreads, riet_rk4_smem_bytes());
                    else        emu_launch(riet_rk4_kernel<false>, R, grid2, kRkThreads, riet_rk4_smem_bytes());
                } else
                if (strict) emu_launch(riet_step_kernel<false, true, kSrcPre, false>, A, grid, g.nthreads, smem);
                else        emu_launch(riet_step_kernel<false, false, kSrcPre, false>, A, grid, g.nthreads, smem);
            } else {
                if (strict) emu_launch(riet_step_kernel<false, true, kSrcPhilox, false>, A, grid, g.nthreads, smem);
                else        emu_launch(riet_step_kernel<false, false, kSrcPhilox, false>, A, grid, g.nthreads, smem);
            }
            if (stats) { // Rietkerk reductions: the stand-alone pass, as tm_measure() runs it
                const int bpt = 3;
                n_records = bpt * n_trials;
                partials.assign(n_records, PartialStats{});
                for (auto &r : partials) r.trial = -1;
                StatsArgs S{g, bufs[cur ^ 1].data(), partials.data(), bpt};
                emu_launch(stats_kernel, S, bpt * n_trials, 64, sizeof(double) * 8 * 64);
            }
        } else {
            std::vector<DpTrial> trials(n_trials);
            for (auto &t : trials) t.dk = dornic_consts(*p, p->a);
            const int bpt = grid_override > 0 ? grid_override : 2;
            n_records = bpt * n_trials;
            partials.assign(n_records, PartialStats{});
            for (auto &r : partials) r.trial = -1;
            DpArgs A{};
            A.g = g;
            A.c.dt = p->dt; A.c.dt2 = p->dt / 2.0; A.c.dt6 = p->dt / 6.0; A.c.b = p->b; A.c.c = p->c;
            A.in = bufs[cur].data(); A.out = bufs[cur ^ 1].data();
            A.trials = trials.data(); A.seed = seed; A.trial0 = trial0; A.step = step0 + s;
            A.replay_u = replay_u; A.stride = stride;
            A.partials = partials.data(); A.counters = counters; A.dornic_out = nullptr; A.blocks_per_trial = bpt;
            const size_t smem = sizeof(double) * 8 * 64;
            if (p->model == TM_MODEL_DP) {
                if (replay_u) emu_launch(dp_step_kernel<kDP, true, true, false>, A, bpt * n_trials, 64, smem);
                else          emu_launch(dp_step_kernel<kDP, false, true, false>, A, bpt * n_trials, 64, smem);
            } else {
                if (replay_u) emu_launch(dp_step_kernel<kDPCubic, true, true, false>, A, bpt * n_trials, 64, smem);
                else          emu_launch(dp_step_kernel<kDPCubic, false, true, false>, A, bpt * n_trials, 64, smem);
            }
        }
        cur ^= 1;
    }
    unpack(g, bufs[cur], fields);
    if (stats_out && nsteps > 0) {
        std::vector<TmStatsDev> sd(n_trials);
        struct FinArgs { const PartialStats *p; int n, nt; double cells; TmStatsDev *out; } F{partials.data(), n_records, n_trials, (double)p->L * p->L, sd.data()};
        auto fin = [](const FinArgs &f) { stats_finalize_kernel(f.p, f.n, f.nt, f.cells, f.out); };
        emu_launch(fin, F, (unsigned)n_trials, 32, 32 * 8 * sizeof(double));
        for (int t = 0; t < n_trials; t++) {
            for (int f = 0; f < 3; f++) { stats_out[8 * t + f] = sd[t].mean[f]; stats_out[8 * t + 3 + f] = (double)sd[t].active[f]; }
            stats_out[8 * t + 6] = (double)sd[t].nonfinite; stats_out[8 * t + 7] = (double)sd[t].negative;
        }
    }
    if (counters_out) { counters_out[0] = counters[0]; counters_out[1] = counters[1]; }
    return 0;
}

void emu_geometry(const tm_params *p, int n_trials, int *out /* ns, TX, nthreads, Lp */)
{
    Geom g{};
    make_geom(*p, n_trials, 0, p->L, &g);
    out[0] = g.ns; out[1] = g.TX; out[2] = g.nthreads; out[3] = g.Lp;
}

// FKE movement of one species through the product's three kernels (csrc/tm_fke.cuh), launch shapes as tm_fke.cu.
// terms: {L, bD, rlo, rhi, clo, chi}; out [L*L]; maxdisp_out [2].
int emu_fke_move(const int *terms, double pref, double eps, const double *rho, const double *gamma, const double *veff,
                 const double *gD, const double *gV, double *out, int *maxdisp_out)
{
    FkeArgs A{};
    A.t.L = terms[0]; A.t.bD = terms[1]; A.t.rlo = terms[2]; A.t.rhi = terms[3]; A.t.clo = terms[4]; A.t.chi = terms[5];
    A.t.pref = pref; A.t.eps = eps;
    const size_t n = (size_t)A.t.L * A.t.L;
    std::vector<double> sA(n), sB(n), sD(n), U1(n), U2(n);
    std::vector<int> disp(n);
    int maxdisp[2] = {0, 0};
    A.rho = rho; A.gamma = gamma; A.veff = veff; A.out = out;
    A.sA = sA.data(); A.sB = sB.data(); A.sD = sD.data(); A.disp = disp.data(); A.maxdisp = maxdisp;
    A.U1 = U1.data(); A.U2 = U2.data(); A.gD = gD; A.gV = gV;
    auto launch = [](auto kernel, const auto &args, unsigned grid, unsigned block, size_t smem) { emu_launch(kernel, args, grid, block, smem); };
    fke_move_enqueue(launch, A);   // the product's launch shapes (tm_fke.cuh)
    maxdisp_out[0] = maxdisp[0]; maxdisp_out[1] = maxdisp[1];
    return 0;
}

// consumer-side operators through the product's kernels (csrc/tm_consumer.cuh), launch shapes as tm_consumer.cu
int emu_cons_gamma(int L, const double *prey, const int *offsets, int eff, double rho_avg_prey, double rho_avg_self, double *gamma)
{
    GammaArgs A{};
    A.L = L; A.eff = eff; A.prey = prey; A.offsets = offsets; A.gamma = gamma;
    A.rho_inverse = rho_avg_prey >= 1.0e-12 ? 1 / rho_avg_prey : 1 / 1.0e-12;
    A.zero = (rho_avg_self < 1.0e-12 || rho_avg_prey < 1.0e-12) ? 1 : 0;
    const int tiles = (L + kConsThreads - 1) / kConsThreads;
    emu_launch(cons_gamma_kernel, A, (unsigned)(tiles * L), kConsThreads, sizeof(int) * 2 * eff);
    return 0;
}

// par = {dt, a, c, gmax, alpha, rW, W0, K1, K2, Dw, Do, A01, H01, E1}; fields [4][L*L] in/out
int emu_cons_rk4(int L, const double *par, double *fields, int nsteps)
{
    const size_t n4 = (size_t)4 * L * L;
    std::vector<double> ya(n4), yb(n4), acc(n4);
    RkStageArgs A{};
    A.L = L;
    A.c.dt = par[0]; A.c.dt2 = par[0] / 2.0; A.c.dt6 = par[0] / 6.0;
    A.c.a = par[1]; A.c.cgmax = par[2] * par[3]; A.c.gmax = par[3]; A.c.alpha = par[4]; A.c.rW = par[5]; A.c.K1 = par[7]; A.c.K2 = par[8];
    A.c.K2W0 = par[8] * par[6]; A.c.Dw = par[9]; A.c.Do = par[10]; A.c.A01 = par[11]; A.c.A01H01 = par[11] * par[12]; A.c.EA1 = par[13] * par[11];
    const int tiles = (L + kConsThreads - 1) / kConsThreads;
    const unsigned grid = (unsigned)(tiles * L);
    for (int s = 0; s < nsteps; s++) {
        A.base = fields; A.acc = acc.data(); A.out = fields;
        A.stage = 1; A.Y = fields; A.Ynext = ya.data(); emu_launch(cons_rk4_stage_kernel, A, grid, kConsThreads, 0);
        A.stage = 2; A.Y = ya.data(); A.Ynext = yb.data(); emu_launch(cons_rk4_stage_kernel, A, grid, kConsThreads, 0);
        A.stage = 3; A.Y = yb.data(); A.Ynext = ya.data(); emu_launch(cons_rk4_stage_kernel, A, grid, kConsThreads, 0);
        A.stage = 4; A.Y = ya.data(); A.Ynext = nullptr; emu_launch(cons_rk4_stage_kernel, A, grid, kConsThreads, 0);
    }
    return 0;
}

// par = {dt, dx, D, v, M, sigma, D0floor}; replay_u == nullptr: Philox
int emu_grazer_dornic(int L, const double *par, unsigned long long seed, unsigned trial, unsigned step, const double *replay_u,
                      unsigned stride, const double *prev, double *gamma, double *out, double *veff, unsigned long long *overflow)
{
    GrazerArgs A{};
    A.L = L; A.dt = par[0]; A.dx = par[1]; A.D = par[2]; A.v = par[3]; A.M = par[4]; A.sigma = par[5]; A.D0floor = par[6];
    grazer_derive(A);
    A.prev = prev; A.gamma = gamma; A.out = out; A.veff = veff;
    A.seed = seed; A.trial = trial; A.step = step; A.replay_u = replay_u; A.stride = stride; A.overflow = overflow;
    *overflow = 0;
    const int tiles = (L + kConsThreads - 1) / kConsThreads;
    if (replay_u) emu_launch(grazer_dornic_kernel<true>, A, (unsigned)(tiles * L), kConsThreads, 0);
    else          emu_launch(grazer_dornic_kernel<false>, A, (unsigned)(tiles * L), kConsThreads, 0);
    return 0;
}

// nsteps SPB = 2 runner steps through the product's launch chain (cons_step2_enqueue), launch shapes as tm_consumer.cu.
// par2 = {dt, dx, d, D0, sigma0, clamp_eps, clamp_lo, gD, gv, gM, gsigma, gD0floor, fr}; cpar as emu_cons_rk4's par.
// strict means of the SPB = 2 / 3 chains (Step2Buffers / Step3Buffers::strict_means) for the emu_cons_step* calls that follow
static int g_strict_means = 0;
void emu_set_strict_means(int on) { g_strict_means = on ? 1 : 0; }

// u_veg/u_graz == nullptr: Philox (seed, trial, step0 + s); else one replay step.  dyn_out[4] = {<P>, <G>, #{P>0}, eff} of the
// state the last step started from.
int emu_cons_step2(int L, const double *par2, const double *cpar, const int *central, int n_central, unsigned long long seed,
                   unsigned trial, unsigned step0, int nsteps, const double *u_veg, const double *u_graz, unsigned stride,
                   double *fields, double *gamma, double *veff, double *dyn_out, unsigned long long *overflow, int rk4_stages)
{
    const size_t n = (size_t)L * L;
    const ConsConsts cc = cons_consts(cpar[0], cpar[1], cpar[2], cpar[3], cpar[4], cpar[5], cpar[6], cpar[7], cpar[8], cpar[9], cpar[10],
                                      cpar[11], cpar[12], cpar[13]);
    const Step2Consts k = step2_consts(L, n_central, par2[12], par2[0], par2[1], par2[2], par2[3], par2[4], par2[5], par2[6], par2[7],
                                       par2[8], par2[9], par2[10], par2[11], cc);
    std::vector<double> star(2 * n), ya(4 * n), yb(4 * n), acc(4 * n), part(3 * kConsReduceBlocks);
    Step2Dyn dyn{};
    Step2Buffers b{};
    b.state = fields; b.star = star.data(); b.ya = ya.data(); b.yb = yb.data(); b.acc = acc.data(); b.gamma = gamma; b.veff = veff;
    b.part = part.data(); b.dyn = &dyn; b.central = central; b.overflow = overflow; b.strict_means = g_strict_means;
    b.replay_veg = u_veg; b.replay_graz = u_graz; b.stride = stride; b.rk4_stages = rk4_stages;
    *overflow = 0;
    auto launch = [](auto kernel, const auto &args, unsigned grid, unsigned block, size_t smem) { emu_launch(kernel, args, grid, block, smem); };
    for (int s = 0; s < nsteps; s++) cons_step2_enqueue(launch, k, b, seed, trial, step0 + (unsigned)s);
    if (b.state != fields) std::memcpy(fields, b.state, sizeof(double) * 4 * n);   // the fused RK4 ping-pongs state and ya
    const double cells = (double)L * (double)L;
    dyn_out[0] = dyn.sum_veg / cells; dyn_out[1] = dyn.sum_graz / cells; dyn_out[2] = dyn.occ_veg; dyn_out[3] = (double)dyn.eff;
    return 0;
}
// The Rietkerk_FPE flavour of the same step (cons_step2_enqueue with k.fpe: the product's own orchestration).  fke = {bound_D, v_row_lo,
// v_row_hi, v_col_lo, v_col_hi}; pref, eps: mu_vel_prefactor, epsilon; gD, gV: the grazer's stencils; M1: its mortality.  The clamp
// constants are par2[5], par2[6] (1e-10 and 0 for the reference's runner).  veff is read, not written.
int emu_cons_step2_fpe(int L, const double *par2, const double *cpar, const int *central, int n_central, unsigned long long seed,
                       unsigned trial, unsigned step0, int nsteps, const double *u_veg, unsigned stride, const int *fke, double pref,
                       double eps, const double *gD, const double *gV, double M1, double *fields, double *gamma, double *veff,
                       double *dyn_out, unsigned long long *overflow, int rk4_stages)
{
    const size_t n = (size_t)L * L;
    ConsConsts cc = cons_consts(cpar[0], cpar[1], cpar[2], cpar[3], cpar[4], cpar[5], cpar[6], cpar[7], cpar[8], cpar[9], cpar[10],
                                cpar[11], cpar[12], cpar[13]);
    cc.fpe = 1; cc.M1 = M1;
    Step2Consts k = step2_consts(L, n_central, par2[12], par2[0], par2[1], par2[2], par2[3], par2[4], par2[5], par2[6], par2[7],
                                 par2[8], par2[9], par2[10], par2[11], cc);
    k.fpe = 1;
    k.fke.L = L; k.fke.bD = fke[0]; k.fke.rlo = fke[1]; k.fke.rhi = fke[2]; k.fke.clo = fke[3]; k.fke.chi = fke[4]; k.fke.pref = pref; k.fke.eps = eps;
    std::vector<double> star(2 * n), ya(4 * n), yb(4 * n), acc(4 * n), part(3 * kConsReduceBlocks);
    Step2Dyn dyn{};
    Step2Buffers b{};
    b.state = fields; b.star = star.data(); b.ya = ya.data(); b.yb = yb.data(); b.acc = acc.data(); b.gamma = gamma; b.veff = veff;
    b.part = part.data(); b.dyn = &dyn; b.central = central; b.overflow = overflow; b.strict_means = g_strict_means;
    b.replay_veg = u_veg; b.replay_graz = u_veg; b.stride = stride; b.rk4_stages = rk4_stages;   // the grazer draws nothing in this flavour
    b.fke_gD = gD; b.fke_gV = gV;
    *overflow = 0;
    auto launch = [](auto kernel, const auto &args, unsigned grid, unsigned block, size_t smem) { emu_launch(kernel, args, grid, block, smem); };
    for (int s = 0; s < nsteps; s++) cons_step2_enqueue(launch, k, b, seed, trial, step0 + (unsigned)s);
    if (b.state != fields) std::memcpy(fields, b.state, sizeof(double) * 4 * n);
    const double cells = (double)L * (double)L;
    dyn_out[0] = dyn.sum_veg / cells; dyn_out[1] = dyn.sum_graz / cells; dyn_out[2] = dyn.occ_veg; dyn_out[3] = (double)dyn.eff;
    return 0;
}
// The SPB = 3 (vegetation + grazer + predator) step through the product's launch chain (csrc/tm_consumer3.cuh).
// par3 = {dt, dx, d, D0, sigma0, clamp_eps, clamp_lo, D1, D2, v1, v2, M1, M2, sigma1, sigma2, D0floor, fr1, fr2};
// cpar = {dt, a, c, gmax, alpha, rW, W0, K1, K2, Dw, Do, A01, H01, E1, A12, H12, E2}; fields [5][L*L], gamma [2][L*L], veff [2][L*L][2];
// u_* == nullptr: Philox (seed, trial, step0 + s), else one replay step; dyn_out[6] = {<P>, <G>, <Pr>, #{P>0}, eff1, eff2}.
int emu_cons_step3(int L, const double *par3, const double *cpar, const int *central, int n_central, unsigned long long seed,
                   unsigned trial, unsigned step0, int nsteps, const double *u_veg, const double *u_graz, const double *u_pred,
                   unsigned stride, double *fields, double *gamma, double *veff, double *dyn_out, unsigned long long *overflow)
{
    const size_t n = (size_t)L * L;
    const Cons3Consts cc = cons3_consts(cpar[0], cpar[1], cpar[2], cpar[3], cpar[4], cpar[5], cpar[6], cpar[7], cpar[8], cpar[9], cpar[10],
                                        cpar[11], cpar[12], cpar[13], cpar[14], cpar[15], cpar[16]);
    const double gD[2] = {par3[7], par3[8]}, gv[2] = {par3[9], par3[10]}, gM[2] = {par3[11], par3[12]}, gs[2] = {par3[13], par3[14]};
    const Step3Consts k = step3_consts(L, n_central, par3[16], par3[17], par3[0], par3[1], par3[2], par3[3], par3[4], par3[5], par3[6],
                                       gD, gv, gM, gs, par3[15], cc);
    std::vector<double> star(3 * n), ya(5 * n), yb(5 * n), acc(5 * n), part(4 * kConsReduceBlocks);
    Step3Dyn dyn{};
    Step3Buffers b{};
    b.state = fields; b.star = star.data(); b.ya = ya.data(); b.yb = yb.data(); b.acc = acc.data(); b.gamma = gamma; b.veff = veff;
    b.part = part.data(); b.dyn = &dyn; b.central = central; b.overflow = overflow; b.strict_means = g_strict_means;
    b.replay[0] = u_veg; b.replay[1] = u_graz; b.replay[2] = u_pred; b.stride = stride;
    *overflow = 0;
    auto launch = [](auto kernel, const auto &args, unsigned grid, unsigned block, size_t smem) { emu_launch(kernel, args, grid, block, smem); };
    for (int s = 0; s < nsteps; s++) cons_step3_enqueue(launch, k, b, seed, trial, step0 + (unsigned)s);
    const double cells = (double)L * (double)L;
    for (int s = 0; s < 3; s++) dyn_out[s] = dyn.sum[s] / cells;
    dyn_out[3] = dyn.occ_veg; dyn_out[4] = (double)dyn.eff[0]; dyn_out[5] = (double)dyn.eff[1];
    return 0;
}
// The Rietkerk_FPE flavour of the SPB = 3 step (cons_step3_enqueue with k.fpe).  fke[2][5] = {bound_D, v_row_lo, v_row_hi, v_col_lo,
// v_col_hi} of grazer and predator; pref[2], eps; gD/gV: their stencils; Mm[2]: mortalities.  u_veg / u_dir: replay arrays or nullptr.
int emu_cons_step3_fpe(int L, const double *par3, const double *cpar, const int *central, int n_central, unsigned long long seed,
                       unsigned trial, unsigned step0, int nsteps, const double *u_veg, const double *u_dir, unsigned stride,
                       const int *fke, const double *pref, double eps, const double *gD1, const double *gV1, const double *gD2,
                       const double *gV2, const double *Mm, double *fields, double *gamma, double *veff, double *dyn_out,
                       unsigned long long *overflow)
{
    const size_t n = (size_t)L * L;
    Cons3Consts cc = cons3_consts(cpar[0], cpar[1], cpar[2], cpar[3], cpar[4], cpar[5], cpar[6], cpar[7], cpar[8], cpar[9], cpar[10],
                                  cpar[11], cpar[12], cpar[13], cpar[14], cpar[15], cpar[16]);
    cc.fpe = 1; cc.M1 = Mm[0]; cc.M2 = Mm[1];
    const double gDc[2] = {par3[7], par3[8]}, gv[2] = {par3[9], par3[10]}, gM[2] = {par3[11], par3[12]}, gs[2] = {par3[13], par3[14]};
    Step3Consts k = step3_consts(L, n_central, par3[16], par3[17], par3[0], par3[1], par3[2], par3[3], par3[4], par3[5], par3[6],
                                 gDc, gv, gM, gs, par3[15], cc);
    k.fpe = 1;
    for (int s = 0; s < 2; s++) {
        FkeTerms &t = k.fke[s];
        t.L = L; t.bD = fke[5 * s]; t.rlo = fke[5 * s + 1]; t.rhi = fke[5 * s + 2]; t.clo = fke[5 * s + 3]; t.chi = fke[5 * s + 4];
        t.pref = pref[s]; t.eps = eps;
    }
    std::vector<double> star(3 * n), ya(5 * n), yb(5 * n), acc(5 * n), part(4 * kConsReduceBlocks);
    Step3Dyn dyn{};
    Step3Buffers b{};
    b.state = fields; b.star = star.data(); b.ya = ya.data(); b.yb = yb.data(); b.acc = acc.data(); b.gamma = gamma; b.veff = veff;
    b.part = part.data(); b.dyn = &dyn; b.central = central; b.overflow = overflow; b.strict_means = g_strict_means;
    b.replay[0] = u_veg; b.replay[1] = u_dir; b.replay[2] = u_dir; b.stride = stride;
    b.fke_gD[0] = gD1; b.fke_gV[0] = gV1; b.fke_gD[1] = gD2; b.fke_gV[1] = gV2;
    *overflow = 0;
    auto launch = [](auto kernel, const auto &args, unsigned grid, unsigned block, size_t smem) { emu_launch(kernel, args, grid, block, smem); };
    for (int s = 0; s < nsteps; s++) cons_step3_enqueue(launch, k, b, seed, trial, step0 + (unsigned)s);
    const double cells = (double)L * (double)L;
    for (int s = 0; s < 3; s++) dyn_out[s] = dyn.sum[s] / cells;
    dyn_out[3] = dyn.occ_veg; dyn_out[4] = (double)dyn.eff[0]; dyn_out[5] = (double)dyn.eff[1];
    return 0;
}
// nsteps RK4 steps of the noise-off DP path through the product's stage kernel (csrc/tm_dpdet.cuh), in place on rho[L*L]
int emu_dp_noise_off(int L, double a, double b, double D, double dt, double dx, double *rho, int nsteps)
{
    const size_t n = (size_t)L * L;
    std::vector<double> ya(n), yb(n), acc(n);
    auto launch = [](auto kernel, const auto &args, unsigned grid, unsigned block, size_t smem) { emu_launch(kernel, args, grid, block, smem); };
    for (int s = 0; s < nsteps; s++) dp_det_step_enqueue(launch, L, a, b, D, dt, dx, rho, ya.data(), yb.data(), acc.data());
    return 0;
}
}
