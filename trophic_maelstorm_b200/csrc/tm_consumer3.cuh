// tm_consumer3.cuh -- sm_100a kernels and launch chain of the SPB = 3 (vegetation + grazer + predator) time step.
// R/ = simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp.  Fields: [0] = P, [1] = G, [2] = Pr (predator), [3] = W, [4] = O.
//
// The body of the time loop of rietkerk_Dornic_2D_MultiSp compiled with -DSPB=3 (R/:3752-4183):
//   vegetation Dornic (R/:3752-3809)                          grazer_dornic_kernel<.., VEG = true>          (tm_consumer.cuh)
//   means of DRho[0..2], vegetation cover (R/:3813-3824)      cons3_reduce_kernel + cons3_finalize_kernel   -> Step3Dyn in device memory
//   calc_gamma_3Sp_NonRefugia (R/:3153-3268, call R/:3850)    cons3_gamma_kernel: both perception fields in one pass
//   the consumers' Dornic loop, s = 1, 2 (R/:3908-4114)       grazer_dornic_kernel twice (own constants, fields, streams, skip flags)
//   RK4_Integrate_Stochastic_MultiSp / f_2Dor_3Sp (R/:2607-2720), clamp (R/:4176)     cons3_rk4_stage_kernel x 4
// Ten launches per step, no host round trip.  A first correct path like the SPB = 2 one it grew from: one thread per cell on
// unpadded fields; the arithmetic keeps the reference's operation order (the TU is built with --fmad=false), so that given the three
// device-side means every field equals the oracle bit for bit.
#pragma once
#include "tm_consumer.cuh"

namespace tml {

struct Step3Dyn {
    double sum[3], occ_veg;      // sums over DRho[0..2]; #{DRho[0] > 0}
    double rho_inverse[3];       // 1 / max(<rho_s>, 1e-12)                                   R/:3171
    int eff[2];                  // disc offsets summed by grazer / predator                  R/:3199, 3204
    int zero[2];                 // gamma_s = 0 everywhere: <rho_s> or <rho_{s-1}> below 1e-12  R/:3206-3209
    int skip[2];                 // <rho_s> == 0: the species' Dornic loop is skipped          R/:3910 (FPE flavour: always)
    int pred_gone, pad_;         // <Pr> below 1e-12: the grazer's sum is not discounted      R/:3221
    int one[2];                  // FPE flavour: <rho_s> below 1e-12: gamma_s = 1              F/:3430
    int maxdisp[2][2];           // FPE flavour: largest |displacement| of each consumer's movement step, zeroed every step
};

struct Reduce3Args {
    long long n;
    const double *P, *G, *R;     // DRho[0..2]
    double *part;                // [gridDim.x][4]
};

// the SPB = 2 reduction tree (cons_reduce_kernel) with one more field
TMC_G void cons3_reduce_kernel(const Reduce3Args A)
{
    TMC_SMEM(smem_raw);
    double *sh = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x, T = kConsReduceThreads;
    const long long chunk = (A.n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > A.n) hi = A.n;
    double sp = 0.0, sg = 0.0, sr = 0.0, oc = 0.0;
    for (long long i = lo + tid; i < hi; i += T) {
        const double p = A.P[i];
        sp += p; sg += A.G[i]; sr += A.R[i];
        if (p > 0) oc += 1;
    }
    sh[tid] = sp; sh[T + tid] = sg; sh[2 * T + tid] = sr; sh[3 * T + tid] = oc;
    __syncthreads();
    for (int s = T / 2; s > 0; s >>= 1) {
        if (tid < s)
            for (int k = 0; k < 4; k++) sh[k * T + tid] += sh[k * T + tid + s];
        __syncthreads();
    }
    if (tid == 0)
        for (int k = 0; k < 4; k++) A.part[4 * blockIdx.x + k] = sh[k * T];
}

struct Finalize3Args {
    int L, nblocks, n_central;
    double fr1, fr2;             // pR[s] / (dx r_max_effective), R/:3409
    const double *part;
    Step3Dyn *dyn;
    int fpe;                     // Rietkerk_FPE flavour: calc_gamma_vel_3Sp_NonRefugia F/:3371-3483
};

TMC_G void cons3_finalize_kernel(const Finalize3Args A)
{
    TMC_SMEM(smem_raw);
    double *sh = reinterpret_cast<double *>(smem_raw);
    const int lane = threadIdx.x;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = lane; b < A.nblocks; b += 32)
        for (int k = 0; k < 4; k++) acc[k] += A.part[4 * b + k];
    for (int k = 0; k < 4; k++) sh[32 * k + lane] = acc[k];
    __syncthreads();
    if (lane != 0) return;
    for (int k = 0; k < 4; k++) { acc[k] = 0.0; for (int l = 0; l < 32; l++) acc[k] += sh[32 * k + l]; }
    const double eps = 1.0e-12;                                       // R/:3160
    const double cells = (double)A.L * (double)A.L;
    Step3Dyn d;
    double avg[3];
    for (int s = 0; s < 3; s++) { d.sum[s] = acc[s]; avg[s] = acc[s] / cells; d.rho_inverse[s] = avg[s] >= eps ? 1 / avg[s] : 1 / eps; }
    d.occ_veg = acc[3];
    double nR_fac = 1 - acc[3] / (A.L * A.L);                         // R/:3822
    if (nR_fac < 0.35) nR_fac = 0.35;
    int new_length = (int)(nR_fac * nR_fac * A.n_central);            // R/:3162
    if (new_length < 1) new_length = 1;
    const double fr[2] = {A.fr1, A.fr2};
    for (int s = 1; s <= 2; s++) {
        int eff = (int)(fr[s - 1] * fr[s - 1] * new_length);          // R/:3199
        if (eff < 1 && fr[s - 1] > 0) eff = 1;                        // R/:3204
        d.eff[s - 1] = eff;
        d.zero[s - 1] = (fr[s - 1] != 0.0 && (avg[s] < eps || avg[s - 1] < eps)) ? 1 : 0;   // R/:3206-3209
        d.skip[s - 1] = (avg[s] == 0.0) ? 1 : 0;                      // R/:3910
        d.one[s - 1] = 0;
        if (A.fpe) {                                                  // F/:3430-3433: a vanished consumer gets gamma = 1
            d.one[s - 1] = (fr[s - 1] != 0.0 && avg[s] < eps) ? 1 : 0;
            d.zero[s - 1] = (fr[s - 1] != 0.0 && !d.one[s - 1] && avg[s - 1] < eps) ? 1 : 0;
            d.skip[s - 1] = 1;                                        // no consumer has a Dornic update in Rietkerk_FPE: rho* = rho
        }
        d.maxdisp[s - 1][0] = 0; d.maxdisp[s - 1][1] = 0;
    }
    d.pred_gone = avg[2] < eps ? 1 : 0;                               // R/:3221
    d.pad_ = 0;
    *A.dyn = d;
}

struct Gamma3Args {
    int L, n_central;            // n_central: upper bound of the offsets staged in shared memory
    double fr1, fr2;
    const double *veg, *graz, *pred;
    const int *offsets;          // [n_central][2]
    double *gamma1, *gamma2;     // [L*L] each, OUT (assigned: R/:3239, 3248)
    const Step3Dyn *dyn;
    int accumulate;              // FPE flavour: the disc sums are added onto the values the arrays hold (F/:3440-3456), and <rho_s> ~ 0 gives 1
};

// calc_gamma_3Sp_NonRefugia R/:3153-3268: gamma_1 = clamp01((sum_disc1 veg (1 - pred / <pred>)) / <veg> / eff1) (undiscounted when the
// predators are gone), gamma_2 = clamp01((sum_disc2 graz / eff2) / <graz>); sums in the reference's order (the disc list's).
TMC_G void cons3_gamma_kernel(const Gamma3Args A)
{
    const Step3Dyn d = *A.dyn;
    const int g = A.L, tid = threadIdx.x;
    const int tiles = (g + kConsThreads - 1) / kConsThreads;
    const int c_i = (int)(blockIdx.x / tiles), c_j = (int)(blockIdx.x % tiles) * kConsThreads + tid;
    const int emax = d.eff[0] > d.eff[1] ? d.eff[0] : d.eff[1];
    TMC_SMEM(smem_raw);
    int *off = reinterpret_cast<int *>(smem_raw);
    for (int k = tid; k < 2 * emax && k < 2 * A.n_central; k += kConsThreads) off[k] = A.offsets[k];
    __syncthreads();
    if (c_j >= g) return;
    const int i = c_i * g + c_j;
    if (A.fr1 != 0.0) {
        if (d.one[0]) A.gamma1[i] = 1.0;
        else if (d.zero[0]) A.gamma1[i] = 0.0;
        else {
            double sum = A.accumulate ? A.gamma1[i] : 0.0;
            for (int k = 0; k < d.eff[0]; k++) {
                int nx = c_i + off[2 * k], ny = c_j + off[2 * k + 1];
                if (nx < 0 || nx >= g) { nx %= g; if (nx < 0) nx += g; }
                if (ny < 0 || ny >= g) { ny %= g; if (ny < 0) ny += g; }
                const int j = nx * g + ny;
                if (d.pred_gone) sum += A.veg[j];                                      // R/:3225
                else sum += A.veg[j] * (1 - A.pred[j] * d.rho_inverse[2]);             // R/:3232
            }
            double v = (sum * d.rho_inverse[0]) / (d.eff[0]);                          // R/:3239
            if (v > 1) v = 1.0; else if (v < 0) v = 0.0;
            A.gamma1[i] = v;
        }
    }
    if (A.fr2 != 0.0) {
        if (d.one[1]) A.gamma2[i] = 1.0;
        else if (d.zero[1]) A.gamma2[i] = 0.0;
        else {
            double sum = A.accumulate ? A.gamma2[i] : 0.0;
            for (int k = 0; k < d.eff[1]; k++) {
                int nx = c_i + off[2 * k], ny = c_j + off[2 * k + 1];
                if (nx < 0 || nx >= g) { nx %= g; if (nx < 0) nx += g; }
                if (ny < 0 || ny >= g) { ny %= g; if (ny < 0) ny += g; }
                sum += A.graz[nx * g + ny];
            }
            double v = (sum / (d.eff[1])) * d.rho_inverse[1];                          // R/:3248
            if (v > 1) v = 1.0; else if (v < 0) v = 0.0;
            A.gamma2[i] = v;
        }
    }
}

struct Cons3Consts {
    double dt, dt2, dt6;
    double a, cgmax, gmax, alpha, rW, K1, K2, K2W0, Dw, Do;
    double A01, A01H01, EA1, A12, A12H12, EA2;
    double M1, M2;               // FPE flavour: the consumers' mortalities, in the reaction terms (f_DorFKE_3Sp F/:2636, 2639)
    int fpe, pad_;
};

inline Cons3Consts cons3_consts(double dt, double a, double c, double gmax, double alpha, double rW, double W0, double K1, double K2,
                                double Dw, double Do, double A01, double H01, double E1, double A12, double H12, double E2)
{
    Cons3Consts k{};
    k.dt = dt; k.dt2 = dt / 2.0; k.dt6 = dt / 6.0;                    // R/:171
    k.a = a; k.cgmax = c * gmax; k.gmax = gmax; k.alpha = alpha; k.rW = rW; k.K1 = K1; k.K2 = K2;
    k.K2W0 = K2 * W0; k.Dw = Dw; k.Do = Do;
    k.A01 = A01; k.A01H01 = A01 * H01; k.EA1 = E1 * A01;              // R/:2614-2616
    k.A12 = A12; k.A12H12 = A12 * H12; k.EA2 = E2 * A12;
    return k;
}

// f_2Dor_3Sp R/:2633-2644 at one site, the reference's operation order
TMC_D void cons3_rhs(const Cons3Consts &c, double P, double G, double R, double W, double O, double Wu, double Wd, double Wl, double Wr,
                     double Ou, double Od, double Ol, double Or_, double f[5])
{
    f[0] = c.cgmax * W * P / (W + c.K1) - (c.A01 * P / (1 + c.A01H01 * P)) * G;
    f[1] = (c.EA1 * P / (1 + c.A01H01 * P)) * G - c.A12 * G / (1 + c.A12H12 * G) * R;
    f[2] = (c.EA2 * G / (1 + c.A12H12 * G)) * R;
    if (c.fpe) {                                                      // f_DorFKE_3Sp F/:2636-2639
        f[1] = (c.EA1 * P / (1 + c.A01H01 * P)) * G - c.M1 * G - c.A12 * G / (1 + c.A12H12 * G) * R;
        f[2] = (c.EA2 * G / (1 + c.A12H12 * G)) * R - c.M2 * R;
    }
    f[3] = c.alpha * (P + c.K2W0) / (P + c.K2) * O - c.gmax * W * P / (W + c.K1) - c.rW * W + (c.Dw) * (Wu + Wd + Wl + Wr - 4 * W);
    f[4] = c.a - c.alpha * (P + c.K2W0) / (P + c.K2) * O + (c.Do) * (Ou + Od + Ol + Or_ - 4 * O);
}

struct Rk3StageArgs {
    int L, stage;                // stage 1..4
    Cons3Consts c;
    const double *base;          // Rho_t [5][L*L]: W, O (fields 3, 4) are read here
    const double *base_sp;       // the three rho* of this step, [3][L*L] (fields 0..2 of Rho_t)
    const double *Y;             // stage state [5][L*L] (stage 1: W, O from base, P, G, Pr from base_sp)
    double *Ynext;               // next stage state (stages 1-3)
    double *acc;                 // K1 (+ 2 K2 (+ 2 K3)) [5][L*L]
    double *out;                 // stage 4: the committed state (may alias base)
    double eps, lo;              // clamp R/:4176
    int clamp;                   // stage 4: apply it (the FPE flavour clamps after the movement step instead)
};

// RK4_Integrate_Stochastic_MultiSp R/:2652-2720, one launch per stage (the SPB = 2 stage kernel with one more species)
TMC_G void cons3_rk4_stage_kernel(const Rk3StageArgs A)
{
    const int g = A.L, n = g * g;
    const int tiles = (g + kConsThreads - 1) / kConsThreads;
    const int r = (int)(blockIdx.x / tiles), q = (int)(blockIdx.x % tiles) * kConsThreads + (int)threadIdx.x;
    if (q >= g) return;
    const int i = r * g + q;
    const int up = (r == 0 ? g - 1 : r - 1) * g + q, dn = (r == g - 1 ? 0 : r + 1) * g + q;      // R/:1622-1626
    const int lf = r * g + (q == 0 ? g - 1 : q - 1), rt = r * g + (q == g - 1 ? 0 : q + 1);
    const double *Ysp = A.stage == 1 ? A.base_sp : A.Y;                 // P, G, Pr of the stage state
    const double *Yw = A.stage == 1 ? A.base : A.Y;                     // W, O of the stage state
    const double *W = Yw + 3 * n, *O = Yw + 4 * n;
    double f[5];
    cons3_rhs(A.c, Ysp[i], Ysp[n + i], Ysp[2 * n + i], W[i], O[i], W[up], W[dn], W[lf], W[rt], O[up], O[dn], O[lf], O[rt], f);
    const double h = A.stage == 3 ? A.c.dt : A.c.dt2;
    for (int s = 0; s < 5; s++) {
        const int k = s * n + i;
        const double b = s < 3 ? A.base_sp[k] : A.base[k];
        if (A.stage == 1) { A.acc[k] = f[s]; A.Ynext[k] = b + (h) * f[s]; }
        else if (A.stage < 4) { A.acc[k] = A.acc[k] + 2.0 * f[s]; A.Ynext[k] = b + (h) * f[s]; }
        else {
            double v = b + (A.c.dt6) * (A.acc[k] + f[s]);               // R/:2691
            if (A.clamp && v < A.eps && v >= A.lo) v = 0.0;             // R/:4176
            A.out[k] = v;
        }
    }
}

// ---- one SPB = 3 runner step as a chain of launches (host side; shared by tm_consumer3.cu and tests/emu) -----------------------
struct Step3Consts {
    int L, n_central;
    double fr1, fr2;             // pR[1], pR[2] over (dx r_max_effective)
    DornicConsts veg;            // R/:3461-3463
    double dt, dx;
    double D[2], v[2], M[2], sigma[2];   // grazer, predator (GrazerArgs)
    double D0floor;              // diff_coefficient[0], R/:4014
    Cons3Consts c;
    double eps, lo;              // clamp R/:4176 (FPE flavour: 1e-10 and 0, F/:4458)
    int fpe;                     // 1: the time-loop body of rietkerk_DorFPE_2D_MultiSp (Rietkerk_FPE, -DSPB=3), see cons_step3_enqueue
    FkeTerms fke[2];             // FPE flavour: movement windows of grazer and predator (set_global_FKE_params F/:195)
};

// FPE flavour: the advection directions of a site, calc_gamma_vel_3Sp_NonRefugia F/:3421-3424: one uniform per species in the order of
// decreasing perception radius -- the vegetation's draw included -- from the cell's direction stream (Philox key: seed xor kDirSeedSalt)
struct Dir3Args {
    int L;
    double fr1, fr2, v1, v2;
    double *veff;                // [2][L*L][2]
    unsigned long long seed; unsigned int trial, step;
    const double *replay_u; unsigned int stride;
    unsigned long long *overflow;
};
constexpr unsigned long long kDirSeedSalt = 0xA0761D6478BD642Full;

template <bool REPLAY>
TMC_G void cons3_direction_kernel(const Dir3Args A)
{
    const long long n = (long long)A.L * A.L;
    const long long i = (long long)blockIdx.x * kConsThreads + threadIdx.x;
    if (i >= n) return;
    typename std::conditional<REPLAY, ReplaySrc, PhiloxSrc>::type src;
    if constexpr (REPLAY) src.begin(A.replay_u + (size_t)i * A.stride, A.stride);
    else src.begin((uint32_t)i, A.trial, A.step, A.seed);
    const double fr[3] = {0.0, A.fr1, A.fr2}, vel[3] = {0.0, A.v1, A.v2};
    int order[3] = {0, 1, 2};
    for (int a = 0; a < 3; a++)
        for (int b = a + 1; b < 3; b++)
            if (fr[order[b]] > fr[order[a]]) { const int t = order[a]; order[a] = order[b]; order[b] = t; }
    for (int o = 0; o < 3; o++) {
        const int s = order[o];
        const double ran = src.next();
        if (s >= 1) {
            double sn, cs;
            sincos_turn128(128.0 * ran, sincos_table(), sn, cs);
            A.veff[2 * ((s - 1) * n + i)] = vel[s] * cs; A.veff[2 * ((s - 1) * n + i) + 1] = vel[s] * sn;
        }
    }
    if (src.overflow_count()) atomicAdd(A.overflow, (unsigned long long)src.overflow_count());
}

// FPE flavour, last launch of a step: the moved consumers replace fields 1 and 2, then `v < eps && v >= lo -> 0` on all five (F/:4458)
struct Fpe3CommitArgs {
    long long n;
    double *state;               // [5][n]
    const double *moved;         // [2][n]
    double eps, lo;
};

TMC_G void cons3_fpe_commit_kernel(const Fpe3CommitArgs A)
{
    const long long i = (long long)blockIdx.x * kConsThreads + threadIdx.x;
    if (i >= A.n) return;
    for (int s = 0; s < 5; s++) {
        double v = (s == 1 || s == 2) ? A.moved[(s - 1) * A.n + i] : A.state[s * A.n + i];
        if (v < A.eps && v >= A.lo) v = 0.0;
        A.state[s * A.n + i] = v;
    }
}

struct Step3Buffers {
    double *state;               // [5][n] P, G, Pr, W, O
    double *star;                // [3][n] rho* of P, G, Pr
    double *ya, *yb, *acc;       // [5][n] each
    double *gamma, *veff;        // [2][n], [2][n][2]
    double *part;                // [kConsReduceBlocks][4]
    Step3Dyn *dyn;
    const int *central;
    unsigned long long *overflow;
    const double *replay[3];     // all null: Philox; else [n][stride] for vegetation, grazer, predator (FPE flavour: [1] = directions)
    unsigned int stride;
    const double *fke_gD[2], *fke_gV[2];   // FPE flavour: the consumers' Gaussian stencils (device)
    int strict_means;            // 1: the three means by cons_seqsum_kernel (left-to-right sums, tm_consumer.cuh) instead of the tree
};

constexpr unsigned long long kPredatorSeedSalt = 0xD1B54A32D192ED03ull;   // the predator's Philox streams (the grazer's: kGrazerSeedSalt)

inline Step3Consts step3_consts(int L, int n_central, double fr1, double fr2, double dt, double dx, double d, double D0, double sigma0,
                                double clamp_eps, double clamp_lo, const double gD[2], const double gv[2], const double gM[2],
                                const double gsigma[2], double gD0floor, const Cons3Consts &c)
{
    Step3Consts k{};
    k.L = L; k.n_central = n_central; k.fr1 = fr1; k.fr2 = fr2;
    const double dx2 = dx * dx;
    const double beta = -d - 4 * D0 / dx2;                            // R/:3461
    const double lambda_exp = exp(beta * dt);
    k.veg.lambda = 2 * beta / (sigma0 * sigma0 * (lambda_exp - 1.0)); // R/:3462
    k.veg.diffc = D0 / dx2;
    k.veg.sigma2 = sigma0 * sigma0;
    k.veg.sigma2_1 = 1 / (sigma0 * sigma0);
    k.veg.ll = k.veg.lambda * lambda_exp;
    k.dt = dt; k.dx = dx;
    for (int s = 0; s < 2; s++) { k.D[s] = gD[s]; k.v[s] = gv[s]; k.M[s] = gM[s]; k.sigma[s] = gsigma[s]; }
    k.D0floor = gD0floor;
    k.c = c; k.eps = clamp_eps; k.lo = clamp_lo;
    return k;
}

template <class Launch>
inline void cons_step3_enqueue(Launch &&launch, const Step3Consts &k, Step3Buffers &b, unsigned long long seed,
                               unsigned int trial, unsigned int step)
{
    const int L = k.L;
    const long long n = (long long)L * L;
    const int tiles = (L + kConsThreads - 1) / kConsThreads;
    const unsigned grid = (unsigned)((long long)tiles * L);
    const bool replay = b.replay[0] != nullptr;
    GrazerArgs V{};
    V.L = L; V.dt = k.dt; V.dx = k.dx; V.D = k.D[0]; V.v = k.v[0]; V.M = k.M[0]; V.sigma = k.sigma[0]; V.D0floor = k.D0floor;
    grazer_derive(V);
    V.prev = b.state; V.gamma = b.gamma; V.out = b.star; V.veff = b.veff;
    V.seed = seed; V.trial = trial; V.step = step; V.replay_u = b.replay[0]; V.stride = b.stride; V.overflow = b.overflow;
    V.veg = k.veg; V.dyn = nullptr; V.skip_flag = nullptr;
    if (replay) launch(grazer_dornic_kernel<true, true>, V, grid, kConsThreads, 0);
    else        launch(grazer_dornic_kernel<false, true>, V, grid, kConsThreads, 0);
    if (b.strict_means) {
        SeqSumArgs S{n, {b.state, b.state + n, b.state + 2 * n}, 3, b.part, kConsReduceBlocks};
        launch(cons_seqsum_kernel, S, 1, 32, 0);
    } else {
        Reduce3Args R{n, b.state, b.state + n, b.state + 2 * n, b.part};
        launch(cons3_reduce_kernel, R, kConsReduceBlocks, kConsReduceThreads, sizeof(double) * 4 * kConsReduceThreads);
    }
    Finalize3Args F{L, kConsReduceBlocks, k.n_central, k.fr1, k.fr2, b.part, b.dyn, k.fpe};
    launch(cons3_finalize_kernel, F, 1, 32, sizeof(double) * 128);
    if (k.fpe) {                                                        // the new advection vectors of this step, F/:3421-3424
        Dir3Args Dd{};
        Dd.L = L; Dd.fr1 = k.fr1; Dd.fr2 = k.fr2; Dd.v1 = k.v[0]; Dd.v2 = k.v[1]; Dd.veff = b.veff;
        Dd.seed = seed ^ kDirSeedSalt; Dd.trial = trial; Dd.step = step; Dd.replay_u = b.replay[1]; Dd.stride = b.stride; Dd.overflow = b.overflow;
        const unsigned gcell = (unsigned)((n + kConsThreads - 1) / kConsThreads);
        if (replay) launch(cons3_direction_kernel<true>, Dd, gcell, kConsThreads, 0);
        else        launch(cons3_direction_kernel<false>, Dd, gcell, kConsThreads, 0);
    }
    Gamma3Args G{};
    G.L = L; G.n_central = k.n_central; G.fr1 = k.fr1; G.fr2 = k.fr2;
    G.veg = b.state; G.graz = b.state + n; G.pred = b.state + 2 * n; G.offsets = b.central;
    G.gamma1 = b.gamma; G.gamma2 = b.gamma + n; G.dyn = b.dyn; G.accumulate = k.fpe;
    launch(cons3_gamma_kernel, G, grid, kConsThreads, sizeof(int) * 2 * k.n_central);
    const unsigned long long salt[2] = {kGrazerSeedSalt, kPredatorSeedSalt};
    for (int s = 0; s < 2; s++) {                                       // the consumers' Dornic loop, R/:3908-4114, s + 1 = 1, 2
        GrazerArgs Z = V;
        Z.D = k.D[s]; Z.v = k.v[s]; Z.M = k.M[s]; Z.sigma = k.sigma[s];
        grazer_derive(Z);
        Z.prev = b.state + (s + 1) * n; Z.out = b.star + (s + 1) * n; Z.gamma = b.gamma + s * n; Z.veff = b.veff + 2 * s * n;
        Z.replay_u = b.replay[s + 1]; Z.dyn = nullptr; Z.skip_flag = &b.dyn->skip[s];
        Z.seed = seed ^ salt[s];
        if (k.fpe) Z.replay_u = b.replay[0];                            // never read: the FPE flavour forces the skip path (rho* = rho)
        if (replay) launch(grazer_dornic_kernel<true, false>, Z, grid, kConsThreads, 0);
        else        launch(grazer_dornic_kernel<false, false>, Z, grid, kConsThreads, 0);
    }
    Rk3StageArgs A{};
    A.L = L; A.c = k.c; A.base = b.state; A.base_sp = b.star; A.acc = b.acc; A.out = b.state; A.eps = k.eps; A.lo = k.lo;
    A.clamp = k.fpe ? 0 : 1;                                            // Rietkerk_FPE clamps after the movement step (F/:4458)
    A.stage = 1; A.Y = b.state; A.Ynext = b.ya; launch(cons3_rk4_stage_kernel, A, grid, kConsThreads, 0);
    A.stage = 2; A.Y = b.ya; A.Ynext = b.yb; launch(cons3_rk4_stage_kernel, A, grid, kConsThreads, 0);
    A.stage = 3; A.Y = b.yb; A.Ynext = b.ya; launch(cons3_rk4_stage_kernel, A, grid, kConsThreads, 0);
    A.stage = 4; A.Y = b.ya; A.Ynext = nullptr; launch(cons3_rk4_stage_kernel, A, grid, kConsThreads, 0);
    if (!k.fpe) return;
    // advdiff_FKE_MultiSp (F/:4308) on the two consumer fields the RK4 left, one after the other (shared scratch: yb = the three source
    // masses, U1, U2; acc = the two moved fields, then the displacements), then the commit + clamp
    for (int s = 0; s < 2; s++) {
        FkeArgs M{};
        M.t = k.fke[s];
        M.rho = b.state + (s + 1) * n; M.gamma = b.gamma + s * n; M.veff = b.veff + 2 * s * n; M.out = b.acc + s * n;
        M.sA = b.yb; M.sB = b.yb + n; M.sD = b.yb + 2 * n; M.U1 = b.yb + 3 * n; M.U2 = b.yb + 4 * n;
        M.disp = reinterpret_cast<int *>(b.acc + 2 * n); M.maxdisp = b.dyn->maxdisp[s];
        M.gD = b.fke_gD[s]; M.gV = b.fke_gV[s];
        fke_move_enqueue(launch, M);
    }
    Fpe3CommitArgs C{n, b.state, b.acc, k.eps, k.lo};
    launch(cons3_fpe_commit_kernel, C, (unsigned)((n + kConsThreads - 1) / kConsThreads), kConsThreads, 0);
}
inline int step3_launches(int fpe) { return fpe ? 18 : 10; }

} // namespace tml
