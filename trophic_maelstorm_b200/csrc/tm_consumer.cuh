// tm_consumer.cuh -- sm_100a kernels of two consumer-side operators of the SPB = 2 (vegetation + grazer) model
// (first part).  R/ = simulations/Rietkerk_FD/rietkerk_bjork_basic.cpp.
//
//   cons_gamma_kernel      calc_gamma_2Sp_NonRefugia R/:2530-2603: the grazer's perception field gamma = clamp01 of the
//                          disc average of the vegetation around each site, normalised by <vegetation>; the disc is a
//                          prefix of the distance-sorted offset list (computeNeighboringSitesCentral R/:1661).  The sum
//                          starts from the value gamma already holds (R/:2581) and runs in the reference's order, so
//                          results equal the reference bit for bit.  A CTA = one lattice row x 128 columns; the offset
//                          list sits in shared memory; prey reads are coalesced along the row and served by L1/L2
//                          (each prey element is read by |disc| sites: 81 for the SPB=2 driver's radius).
//   grazer_dornic_kernel   the grazer's Dornic update with upwind advection, R/:3908-4114: per cell one uniform for the
//                          advection direction (drawn for every cell, before the skip test), the advection-corrected
//                          diffusion (R/:4008-4018), alpha_i with the upwind neighbours (R/:4044), per-cell beta/lambda
//                          (an exp per cell, R/:4057-4059), then the same Poisson/Gamma ladder with Gaussian shortcuts as
//                          the vegetation (dornic_cell<kRietkerk>).  One thread per cell; Philox stream keyed by
//                          (seed, trial, step, cell) or host-supplied uniforms (replay); bit-identical to the oracle.
//   cons_rk4_stage_kernel  RK4_Integrate_Stochastic_2Sp R/:2472-2527 with f_2Dor_2Sp R/:2445-2470 over (P, G, W, O): one
//                          launch per stage.  Stage k reads the stage state Y (5-point neighbours of W and O, periodic),
//                          writes Y' = Rho_t + h K_k and the running K1 + 2 K2 + 2 K3 (+ K4) in the reference's left-to-
//                          right order; the last stage commits Rho_t += dt/6 (...).  HBM-bound: per stage 4 fields read
//                          (+ base state) and 8 written.  STRICT keeps the reference's operation order (no contraction:
//                          the TU is built with --fmad=false), bit-identical to the oracle.
#pragma once
#include <stdint.h>
#include "tm_samplers.cuh"
#include "tm_fke.cuh"      // the FPE flavour of the step moves the grazer with the FKE kernels
#include <type_traits>
#include <math.h>

#if defined(__CUDACC__)
#define TMC_D __device__ __forceinline__
#define TMC_G static __global__   // static: tm_consumer.cu and tm_consumer3.cu each compile their own copy
#define TMC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#include "tm_cpu_emu.h" // tests/emu only
#define TMC_D inline
#define TMC_G
#define TMC_SMEM(name) unsigned char *name = tm_emu_dyn_smem()
#endif

namespace tml {

constexpr int kConsThreads = 128;

// Scalars of one SPB = 2 runner step that the reference derives on the host between its loops (R/:3813-3824, 2533-2578,
// 3910).  Here a two-kernel reduction leaves them in device memory and the following kernels read them there, so a
// whole time step is enqueued without a host round trip.
struct Step2Dyn {
    double sum_veg, sum_graz, occ_veg;   // sums over DRho[0], DRho[1]; #{DRho[0] > 0}
    double rho_inverse;                  // 1 / max(<veg>, 1e-12)                          R/:2548
    int eff;                             // disc offsets summed: int(fr^2 * new_length)    R/:2539, 2569
    int zero;                            // <grazer> or <veg> below 1e-12: gamma = 0       R/:2575-2578
    int skip_graz;                       // <grazer> == 0: the grazer's Dornic loop is skipped, R/:3910 (FPE flavour: always)
    int one;                             // FPE flavour: <grazer> below 1e-12: gamma = 1           F/:2579
    int maxdisp[2];                      // FPE flavour: largest |displacement| of the movement step (fke_prepare_kernel), zeroed every step
};

constexpr int kConsReduceThreads = 256;
constexpr int kConsReduceBlocks = 296;   // two CTAs per SM; the summation tree below is fixed by these two constants

struct ReduceArgs {
    long long n;                 // L*L
    const double *P, *G;         // DRho[0], DRho[1]
    double *part;                // [gridDim.x][3]
};

// mean_of_vector / occupied_sites_of_vector (R/:1211, 1245) of the vegetation and the grazer as block partials.  The
// reference sums left to right; this is a fixed tree (thread-strided partials of a contiguous chunk, pairwise in shared
// memory, then block partials in order), reproducible run to run and machine to machine but rounded differently from the
// sequential sum: <rho> differs by a few ulp (tests: <= 1e-13 relative), the integer count is exact.
TMC_G void cons_reduce_kernel(const ReduceArgs A)
{
    TMC_SMEM(smem_raw);
    double *sh = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x, T = kConsReduceThreads;
    const long long chunk = (A.n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > A.n) hi = A.n;
    double sp = 0.0, sg = 0.0, oc = 0.0;
    for (long long i = lo + tid; i < hi; i += T) {
        const double p = A.P[i];
        sp += p; sg += A.G[i];
        if (p > 0) oc += 1;
    }
    sh[tid] = sp; sh[T + tid] = sg; sh[2 * T + tid] = oc;
    __syncthreads();
    for (int s = T / 2; s > 0; s >>= 1) {
        if (tid < s) { sh[tid] += sh[tid + s]; sh[T + tid] += sh[T + tid + s]; sh[2 * T + tid] += sh[2 * T + tid + s]; }
        __syncthreads();
    }
    if (tid == 0) { A.part[3 * blockIdx.x] = sh[0]; A.part[3 * blockIdx.x + 1] = sh[T]; A.part[3 * blockIdx.x + 2] = sh[2 * T]; }
}

// Strict means (tm_cons_step2_set_strict_means / tm_cons3_set_strict_means): mean_of_vector's own left-to-right sum (R/:1211-1219) --
// ONE thread per field walks it in index order, a third/fourth lane counts the vegetation cover -- written as block 0's partial with
// every other block's partial zeroed, so that the finalize kernel's additions (x + 0.0) hand the sequential sums on unchanged and the
// whole step equals the reference-pinned oracle bit for bit WITHOUT being given its means.  A dependent DADD chain of L*L terms: for
// parity runs, not for production (70 ms at L = 4096 against 3 ms for the step).
struct SeqSumArgs {
    long long n;
    const double *f[3];          // the biotic fields DRho[0 .. nf-1]
    int nf;                      // 2 (SPB = 2) or 3 (SPB = 3)
    double *part;                // [nblocks][nf + 1]
    int nblocks;
};

TMC_G void cons_seqsum_kernel(const SeqSumArgs A)
{
    const int lane = threadIdx.x, w = A.nf + 1;
    for (long long k = w + lane; k < (long long)A.nblocks * w; k += blockDim.x) A.part[k] = 0.0;
    if (lane < A.nf) {
        const double *x = A.f[lane];
        double s = 0.0;
        for (long long i = 0; i < A.n; ++i) s += x[i];
        A.part[lane] = s;
    } else if (lane == A.nf) {
        const double *x = A.f[0];
        double oc = 0.0;
        for (long long i = 0; i < A.n; ++i) if (x[i] > 0) oc += 1;
        A.part[lane] = oc;
    }
}

struct FinalizeArgs {
    int L, nblocks, n_central;   // n_central: length of the full disc list of computeNeighboringSitesCentral(int(r_max))
    double fr;                   // pR[1] / (dx r_max_effective), R/:3409
    const double *part;
    Step2Dyn *dyn;
    int fpe;                     // Rietkerk_FPE flavour: calc_gamma_vel_2Sp_NonRefugia F/:2534-2607
};

TMC_G void cons_finalize_kernel(const FinalizeArgs A)
{
    // 32 threads: lane l sums partials l, l+32, ... in order, then the 32 lane sums are added in lane order by lane 0
    // (a fixed tree again; through shared memory so that the host-thread emulation runs the same code)
    TMC_SMEM(smem_raw);
    double *sh = reinterpret_cast<double *>(smem_raw);
    const int lane = threadIdx.x;
    double sp = 0.0, sg = 0.0, oc = 0.0;
    for (int b = lane; b < A.nblocks; b += 32) { sp += A.part[3 * b]; sg += A.part[3 * b + 1]; oc += A.part[3 * b + 2]; }
    sh[lane] = sp; sh[32 + lane] = sg; sh[64 + lane] = oc;
    __syncthreads();
    if (lane != 0) return;
    sp = 0.0; sg = 0.0; oc = 0.0;
    for (int l = 0; l < 32; l++) { sp += sh[l]; sg += sh[32 + l]; oc += sh[64 + l]; }
    const double eps = 1.0e-12;                                       // R/:2537
    const double cells = (double)A.L * (double)A.L;
    const double avg_veg = sp / cells, avg_graz = sg / cells;        // R/:3813-3818
    double nR_fac = 1 - oc / (A.L * A.L);                             // R/:3822
    if (nR_fac < 0.35) nR_fac = 0.35;
    int new_length = (int)(nR_fac * nR_fac * A.n_central);            // R/:2539
    if (new_length < 1 && !A.fpe) new_length = 1;                     // R/:2540; the FPE twin has no such line (F/:2543)
    Step2Dyn d;
    d.sum_veg = sp; d.sum_graz = sg; d.occ_veg = oc;
    d.eff = (int)(A.fr * A.fr * new_length);                          // R/:2569
    d.rho_inverse = avg_veg >= eps ? 1 / avg_veg : 1 / eps;           // R/:2548
    d.zero = (avg_graz < eps || avg_veg < eps) ? 1 : 0;               // R/:2575-2578
    d.skip_graz = (avg_graz == 0.0) ? 1 : 0;                          // R/:3910
    d.one = 0;
    if (A.fpe) {                                                      // F/:2578-2581: a vanished consumer gets gamma = 1
        d.one = avg_graz < eps ? 1 : 0;
        d.zero = (!d.one && avg_veg < eps) ? 1 : 0;
        d.skip_graz = 1;                                              // the grazer has no Dornic update in Rietkerk_FPE: rho* = rho
    }
    d.maxdisp[0] = 0; d.maxdisp[1] = 0;
    *A.dyn = d;
}

struct GammaArgs {
    int L, eff;                  // lattice side; number of disc offsets summed (R/:2569)
    const double *prey;          // [L*L] Rho_t[s-1]
    const int *offsets;          // [eff][2] = {dx (rows), dy (columns)}
    double rho_inverse;          // 1 / max(<prey>, 1e-12)  (R/:2548)
    int zero;                    // <self> or <prey> below 1e-12: gamma = 0 everywhere (R/:2575-2578)
    int one;                     // FPE flavour, <self> below 1e-12: gamma = 1 everywhere (F/:2579)
    double *gamma;               // [L*L] in/out
    const Step2Dyn *dyn;         // non-null: eff, rho_inverse and zero come from the device-side scalars of this step
                                 // (the launch reserves shared memory for `eff` offsets = the upper bound)
};

TMC_G void cons_gamma_kernel(const GammaArgs A_in)
{
    GammaArgs A = A_in;
    if (A.dyn) {
        const Step2Dyn d = *A.dyn;
        if (d.eff < A.eff) A.eff = d.eff;          // d.eff <= fr^2 n_central by construction
        A.rho_inverse = d.rho_inverse; A.zero = d.zero; A.one = d.one;
    }
    const int g = A.L, tid = threadIdx.x;
    const int tiles = (g + kConsThreads - 1) / kConsThreads;
    const int c_i = (int)(blockIdx.x / tiles), c_j = (int)(blockIdx.x % tiles) * kConsThreads + tid;
    TMC_SMEM(smem_raw);
    int *off = reinterpret_cast<int *>(smem_raw);
    for (int k = tid; k < 2 * A.eff; k += kConsThreads) off[k] = A.offsets[k];
    __syncthreads();
    if (c_j >= g) return;
    const int i = c_i * g + c_j;
    if (A.one) { A.gamma[i] = 1.0; return; }
    if (A.zero) { A.gamma[i] = 0.0; return; }
    double acc = A.gamma[i];
    for (int k = 0; k < A.eff; k++) {
        // (c + d + L) % L of R/:1700-1701 for |d| <= L; larger discs wrap more than once
        int nx = c_i + off[2 * k], ny = c_j + off[2 * k + 1];
        if (nx < 0 || nx >= g) { nx %= g; if (nx < 0) nx += g; }
        if (ny < 0 || ny >= g) { ny %= g; if (ny < 0) ny += g; }
        acc += A.prey[nx * g + ny];
    }
    acc = (acc * A.rho_inverse) / (A.eff);
    if (acc > 1) acc = 1.0; else if (acc < 0) acc = 0.0;
    A.gamma[i] = acc;
}

struct GrazerArgs {
    int L;
    double dt, dx, D, v, M, sigma, D0floor;   // D[s], v[s], M[s], sigma[s]; diff_coefficient[0] (R/:4014)
    double dx1, dx1_2, dxb2, sigma2_1;        // 1/dx, 1/dx^2, dx/2, 1/sigma^2 (R/:3455, 3463): set by grazer_derive() on the host --
                                              // the same IEEE operations the kernel used to repeat in every thread
    const double *prev;                       // DRho[s] [L*L]
    double *gamma;                            // [L*L] in/out (clamped at 1, R/:3915)
    double *out, *veff;                       // rho* [L*L]; advection vectors [L*L][2]
    unsigned long long seed; unsigned int trial, step;
    const double *replay_u; unsigned int stride;   // REPLAY: [L*L][stride]
    unsigned long long *overflow;
    DornicConsts veg;                         // VEG variant: the vegetation's constants (R/:3461-3463), same for every cell
    const Step2Dyn *dyn;                      // non-null and dyn->skip_graz: rho* = prev, nothing drawn (R/:3910)
    const int *skip_flag;                     // non-null and *skip_flag != 0: the same (the SPB = 3 step keeps one flag per consumer)
};

inline void grazer_derive(GrazerArgs &A)
{
    const double dx2 = A.dx * A.dx;
    A.dx1 = 1 / A.dx; A.dx1_2 = 1 / (dx2); A.dxb2 = A.dx / 2.0;                              // R/:3455
    A.sigma2_1 = 1 / (A.sigma * A.sigma);                                                     // R/:3463
}

// VEG = true: the vegetation's Dornic loop R/:3752-3809 on an unpadded field (no advection, no direction draw, constant
// beta/lambda, alpha_i = (D0/dx^2) (up + down + left + right) in the reference's order): the first kernel of the SPB = 2 step.
template <bool REPLAY, bool VEG = false>
TMC_G void grazer_dornic_kernel(const GrazerArgs A)
{
    const int g = A.L;
    const int tiles = (g + kConsThreads - 1) / kConsThreads;
    const int r = (int)(blockIdx.x / tiles), q = (int)(blockIdx.x % tiles) * kConsThreads + (int)threadIdx.x;
    if (q >= g) return;
    const int i = r * g + q;
    const int up = (r == 0 ? g - 1 : r - 1), dn = (r == g - 1 ? 0 : r + 1), lf = (q == 0 ? g - 1 : q - 1), rt = (q == g - 1 ? 0 : q + 1);
    const double dx1 = A.dx1, dx1_2 = A.dx1_2, dxb2 = A.dxb2, sigma2_1 = A.sigma2_1;        // grazer_derive()
    typename std::conditional<REPLAY, ReplaySrc, PhiloxSrc>::type src;
    if constexpr (REPLAY) src.begin(A.replay_u + (size_t)i * A.stride, A.stride);
    else src.begin((uint32_t)i, A.trial, A.step, A.seed);
    if constexpr (VEG) {
        const double *P = A.prev;
        const double alpha_i = A.veg.diffc * (P[up * g + q] + P[dn * g + q] + P[r * g + lf] + P[r * g + rt]);   // R/:3756
        const double rho_i = P[i];
        double res = rho_i;
        if (!(alpha_i == 0 && rho_i == 0)) res = dornic_cell<kRietkerk>(A.veg, src, alpha_i, rho_i);          // R/:3759-3793
        A.out[i] = res;
        if (src.overflow_count()) atomicAdd(A.overflow, (unsigned long long)src.overflow_count());
        return;
    }
    if ((A.dyn && A.dyn->skip_graz) || (A.skip_flag && *A.skip_flag)) { A.out[i] = A.prev[i]; return; }
    double gm = A.gamma[i];
    if (gm > 1) { gm = 1; A.gamma[i] = gm; }
    const double ran = src.next();                                                         // R/:3993
    double sn, cs;
    sincos_turn128(128.0 * ran, sincos_table(), sn, cs);
    const double vx = A.v * cs, vy = A.v * sn;
    A.veff[2 * i] = vx; A.veff[2 * i + 1] = vy;
    const double vx_abs = vx * ((0 < vx) - (vx < 0)), vy_abs = vy * ((0 < vy) - (vy < 0));
    const double gp = (1 - gm);
    double dex = (A.D - (dxb2) * (gp * vx_abs * (1 - A.dt * dx1 * gp * vx_abs))) * dx1_2;
    double dey = (A.D - (dxb2) * (gp * vy_abs * (1 - A.dt * dx1 * gp * vy_abs))) * dx1_2;
    if (dex < A.D0floor) dex = A.D0floor;
    if (dey < A.D0floor) dey = A.D0floor;
    const double *P = A.prev;
    const double hx = P[r * g + ((0 < vx) ? lf : rt)], hy = P[((0 < vy) ? up : dn) * g + q];
    const double alpha_i = gm * (dey * (P[up * g + q] + P[dn * g + q]) + dex * (P[r * g + lf] + P[r * g + rt]))
                         + dx1 * (gp) * (vx_abs * hx + vy_abs * hy);
    const double rho_i = P[i];
    double res = rho_i;
    if (!(alpha_i == 0 && rho_i == 0)) {
        const double beta = -A.M - 2 * gm * (dex + dey) - (1 - gm) * (vx_abs + vy_abs) * dx1;
        const double lambda_exp = exp_any((beta) * A.dt);
        DornicConsts k;
        k.lambda = 2 * (beta) * (sigma2_1) / (lambda_exp - 1.0);
        k.ll = k.lambda * lambda_exp; k.sigma2 = A.sigma * A.sigma; k.sigma2_1 = sigma2_1; k.diffc = 0.0;
        res = dornic_cell<kRietkerk>(k, src, alpha_i, rho_i);
    }
    A.out[i] = res;
    if (src.overflow_count()) atomicAdd(A.overflow, (unsigned long long)src.overflow_count());
}

struct ConsConsts {
    double dt, dt2, dt6;
    double a, cgmax, gmax, alpha, rW, K1, K2, K2W0, Dw, Do, A01, A01H01, EA1;
    double M1;                   // FPE flavour: the grazer's mortality, in the reaction terms (f_DorFKE_2Sp F/:2466)
    int fpe, pad_;
};

struct RkStageArgs {
    int L, stage;                // stage 1..4
    ConsConsts c;
    const double *base;          // Rho_t   [4][L*L]  (P, G, W, O)
    const double *base_pg;       // non-null: fields 0, 1 of Rho_t (the two rho* of this step) live here as [2][L*L]
    const double *Y;             // stage state (== base for stage 1)
    double *Ynext;               // next stage state (stages 1-3)
    double *acc;                 // K1 (+ 2 K2 (+ 2 K3)) [4][L*L]
    double *out;                 // stage 4: the committed state (may alias base)
    int clamp;                   // stage 4: apply `if (v < eps && v >= lo) v = 0` (R/:4176) to the committed values
    double eps, lo;
};

// f_2Dor_2Sp R/:2460-2468 at one site, the reference's operation order
TMC_D void cons_rhs(const ConsConsts &c, double P, double G, double W, double O, double Wu, double Wd, double Wl, double Wr,
                    double Ou, double Od, double Ol, double Or_, double f[4])
{
    f[0] = c.cgmax * W * P / (W + c.K1) - (c.A01 * P / (1 + c.A01H01 * P)) * G;
    f[1] = (c.EA1 * P / (1 + c.A01H01 * P)) * G;
    if (c.fpe) f[1] = (c.EA1 * P / (1 + c.A01H01 * P)) * G - c.M1 * G;   // f_DorFKE_2Sp F/:2466
    f[2] = c.alpha * (P + c.K2W0) / (P + c.K2) * O - c.gmax * W * P / (W + c.K1) - c.rW * W + (c.Dw) * (Wu + Wd + Wl + Wr - 4 * W);
    f[3] = c.a - c.alpha * (P + c.K2W0) / (P + c.K2) * O + (c.Do) * (Ou + Od + Ol + Or_ - 4 * O);
}

TMC_G void cons_rk4_stage_kernel(const RkStageArgs A)
{
    const int g = A.L, n = g * g;
    const int tiles = (g + kConsThreads - 1) / kConsThreads;
    const int r = (int)(blockIdx.x / tiles), q = (int)(blockIdx.x % tiles) * kConsThreads + (int)threadIdx.x;
    if (q >= g) return;
    const int i = r * g + q;
    const int up = (r == 0 ? g - 1 : r - 1) * g + q, dn = (r == g - 1 ? 0 : r + 1) * g + q;      // R/:1622-1626
    const int lf = r * g + (q == 0 ? g - 1 : q - 1), rt = r * g + (q == g - 1 ? 0 : q + 1);
    const double *Y = A.Y, *W = Y + 2 * n, *O = Y + 3 * n;
    const double *B01 = A.base_pg ? A.base_pg : A.base;
    const double *Ypg = (A.stage == 1 && A.base_pg) ? A.base_pg : Y;
    double f[4];
    cons_rhs(A.c, Ypg[i], Ypg[n + i], W[i], O[i], W[up], W[dn], W[lf], W[rt], O[up], O[dn], O[lf], O[rt], f);
    const double h = A.stage == 3 ? A.c.dt : A.c.dt2;
    for (int s = 0; s < 4; s++) {
        const int k = s * n + i;
        const double b = s < 2 ? B01[k] : A.base[k];
        if (A.stage == 1) { A.acc[k] = f[s]; A.Ynext[k] = b + (h) * f[s]; }
        else if (A.stage < 4) { A.acc[k] = A.acc[k] + 2.0 * f[s]; A.Ynext[k] = b + (h) * f[s]; }
        else {
            double v = b + (A.c.dt6) * (A.acc[k] + f[s]);     // Rho_t += dt6 (K1 + 2 K2 + 2 K3 + K4), R/:2506
            if (A.clamp && v < A.eps && v >= A.lo) v = 0.0;           // R/:4176
            A.out[k] = v;
        }
    }
}

// All four RK4 stages of RK4_Integrate_Stochastic_2Sp (R/:2472-2527) and the clamp (R/:4176) in ONE launch: a CTA owns a
// kT x kT tile of the lattice and recomputes the RK4 dependency cone around it (three stencil applications: the stage-k
// values are made on a (kT + 8 - 2k)^2 region).  One thread per cell of the (kT+6)^2 stage-1 region keeps its own P, G, base
// values and running K sums in registers; only the W, O stage values -- the two fields with a Laplacian -- pass through two
// ping-pong shared-memory tiles.  Per cell the arithmetic is the stage kernels' (cons_rhs, then `b + h f` / `acc + 2 f` /
// `b + dt6 (acc + f)` in the reference's order), so results are bit-identical to the four-launch form; HBM traffic drops from
// 68 field passes to ~8 (tile halos are L2 hits).
constexpr int kFuseT = 26;                 // tile side
constexpr int kFuseE = kFuseT + 6;         // stage-1 region side (one thread per cell)
constexpr int kFuseS = kFuseT + 8;         // shared tile side (stage-0 W, O incl. the outer ring)
constexpr int kFuseThreads = 1024;         // >= kFuseE^2 = 900 and >= the ring loaders

struct RkFusedArgs {
    int L;
    ConsConsts c;
    const double *base;          // W, O at [2][L*L], [3][L*L] of this [4][L*L] block (P, G there are ignored when base_pg != null)
    const double *base_pg;       // rho* of P and G as [2][L*L] (null: fields 0, 1 of base)
    double *out;                 // [4][L*L]; must NOT alias base: a tile reads W, O of its neighbours' cells
    int clamp; double eps, lo;
};

#if defined(__CUDACC__)
static __global__ void __launch_bounds__(kFuseThreads) cons_rk4_fused_kernel(const RkFusedArgs A)
#else
void cons_rk4_fused_kernel(const RkFusedArgs A)
#endif
{
    TMC_SMEM(smem_raw);
    double *SW0 = reinterpret_cast<double *>(smem_raw), *SO0 = SW0 + kFuseS * kFuseS;
    double *SW1 = SO0 + kFuseS * kFuseS, *SO1 = SW1 + kFuseS * kFuseS;
    const int g = A.L, n = g * g, tid = threadIdx.x;
    const int tiles = (g + kFuseT - 1) / kFuseT;
    const int ty0 = (int)(blockIdx.x / tiles) * kFuseT, tx0 = (int)(blockIdx.x % tiles) * kFuseT;
    auto wrap = [g](int v) { v %= g; if (v < 0) v += g; return v; };
    // stage-0 W, O of the (kT+8)^2 region into shared memory
    for (int k = tid; k < kFuseS * kFuseS; k += kFuseThreads) {
        const int sy = k / kFuseS, sx = k % kFuseS;
        const int i = wrap(ty0 - 4 + sy) * g + wrap(tx0 - 4 + sx);
        SW0[k] = A.base[2 * n + i]; SO0[k] = A.base[3 * n + i];
    }
    const bool mine = tid < kFuseE * kFuseE;
    const int ey = mine ? tid / kFuseE : 0, ex = mine ? tid % kFuseE : 0;       // position in the stage-1 region
    const int sc = (ey + 1) * kFuseS + (ex + 1);                                 // ... and in the shared tiles
    const int gy = wrap(ty0 - 3 + ey), gx = wrap(tx0 - 3 + ex);
    const int i = gy * g + gx;
    const double *B01 = A.base_pg ? A.base_pg : A.base;
    const double bP = B01[i], bG = B01[n + i];
    __syncthreads();
    const double bW = SW0[sc], bO = SO0[sc];
    double yP = bP, yG = bG, yW = bW, yO = bO, acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int stage = 1; stage <= 4; stage++) {       // unrolled: the ping-pong tiles, region tests and step sizes resolve statically
        const double *RW = (stage & 1) ? SW0 : SW1, *RO = (stage & 1) ? SO0 : SO1;   // read tiles of this stage
        double *WW = (stage & 1) ? SW1 : SW0, *WO_ = (stage & 1) ? SO1 : SO0;        // ... written for the next
        const int m = stage - 1;                                                     // region shrinks by one ring per stage
        const bool in = mine && ey >= m && ey < kFuseE - m && ex >= m && ex < kFuseE - m;
        if (in) {
            double f[4];
            cons_rhs(A.c, yP, yG, yW, yO, RW[sc - kFuseS], RW[sc + kFuseS], RW[sc - 1], RW[sc + 1],
                     RO[sc - kFuseS], RO[sc + kFuseS], RO[sc - 1], RO[sc + 1], f);
            const double h = stage == 3 ? A.c.dt : A.c.dt2;
            const double b[4] = {bP, bG, bW, bO};
            if (stage < 4) {
                double y[4];
                for (int s = 0; s < 4; s++) {
                    acc[s] = stage == 1 ? f[s] : acc[s] + 2.0 * f[s];
                    y[s] = b[s] + (h) * f[s];
                }
                yP = y[0]; yG = y[1]; yW = y[2]; yO = y[3];
                WW[sc] = yW; WO_[sc] = yO;
            } else if (ty0 + (ey - 3) < g && tx0 + (ex - 3) < g) {                   // the tile's own cells inside the lattice
                for (int s = 0; s < 4; s++) {
                    double v = b[s] + (A.c.dt6) * (acc[s] + f[s]);                   // R/:2506
                    if (A.clamp && v < A.eps && v >= A.lo) v = 0.0;                  // R/:4176
                    A.out[s * n + i] = v;
                }
            }
        }
        __syncthreads();
    }
}

// ---- one SPB = 2 runner step as a chain of launches (host side; shared by tm_consumer.cu and tests/emu) ----------------
struct Step2Consts {
    int L, n_central;            // n_central: length of the full disc list (the gamma kernel's shared-memory bound)
    double fr;                   // pR[1] / (dx r_max_effective)
    DornicConsts veg;            // R/:3461-3463
    double dt, dx, D, v, M, sigma, D0floor;   // the grazer's (GrazerArgs)
    ConsConsts c;
    double eps, lo;              // clamp R/:4176 (FPE flavour: 1e-10 and 0, F/:4458)
    int fpe;                     // 0: rietkerk_Dornic_2D_MultiSp (Rietkerk_FD); 1: rietkerk_DorFPE_2D_MultiSp (Rietkerk_FPE), below
    FkeTerms fke;                // FPE flavour: the grazer's movement windows (set_global_FKE_params F/:195)
};

struct Step2Buffers {
    double *state;               // [4][n] P, G, W, O: Rho_dt == DRho between steps
    double *star;                // [2][n] rho* of P and G
    double *ya, *yb, *acc;       // [4][n] each
    double *gamma, *veff;        // [n], [n][2]
    double *part;                // [kConsReduceBlocks][3]
    Step2Dyn *dyn;
    const int *central;          // [n_central][2]
    unsigned long long *overflow;
    const double *replay_veg, *replay_graz;   // both null: Philox; else [n][stride] each
    unsigned int stride;
    int rk4_stages;              // 1: the four-launch RK4 (kept as the cross-check); 0: cons_rk4_fused_kernel, which writes the
                                 // new state into `ya` -- cons_step2_enqueue then swaps `state` and `ya`
    const double *fke_gD, *fke_gV;   // FPE flavour: the grazer's two Gaussian stencils (device)
    int strict_means;            // 1: <P>, <G> by cons_seqsum_kernel (the reference's left-to-right sums) instead of the reduction tree
};

// FPE flavour, last launch of a step: the moved grazer field replaces field 1, then `v < eps && v >= lo -> 0` on all four (F/:4458)
struct FpeCommitArgs {
    long long n;
    double *state;               // [4][n]
    const double *moved;         // [n]
    double eps, lo;
};

TMC_G void cons_fpe_commit_kernel(const FpeCommitArgs A)
{
    const long long i = (long long)blockIdx.x * kConsThreads + threadIdx.x;
    if (i >= A.n) return;
    for (int s = 0; s < 4; s++) {
        double v = s == 1 ? A.moved[i] : A.state[s * A.n + i];
        if (v < A.eps && v >= A.lo) v = 0.0;
        A.state[s * A.n + i] = v;
    }
}

// The grazer's Philox streams use the seed with this constant xor-ed in, so that the two species of a cell draw from
// different streams (the vegetation keeps the plain seed: its rho* equals the SPB = 1 step's for the same seed).
constexpr unsigned long long kGrazerSeedSalt = 0x9E3779B97F4A7C15ull;

// host-side constants of a step from the runner's arguments (R/:3455-3480), shared by the C ABI and the emulator
inline Step2Consts step2_consts(int L, int n_central, double fr, double dt, double dx, double d, double D0, double sigma0,
                                double clamp_eps, double clamp_lo, double gD, double gv, double gM, double gsigma, double gD0floor,
                                const ConsConsts &c)
{
    Step2Consts k{};
    k.L = L; k.n_central = n_central; k.fr = fr;
    const double dx2 = dx * dx;
    const double beta = -d - 4 * D0 / dx2;                            // R/:3461
    const double lambda_exp = exp(beta * dt);
    k.veg.lambda = 2 * beta / (sigma0 * sigma0 * (lambda_exp - 1.0)); // R/:3462
    k.veg.diffc = D0 / dx2;
    k.veg.sigma2 = sigma0 * sigma0;
    k.veg.sigma2_1 = 1 / (sigma0 * sigma0);
    k.veg.ll = k.veg.lambda * lambda_exp;
    k.dt = dt; k.dx = dx; k.D = gD; k.v = gv; k.M = gM; k.sigma = gsigma; k.D0floor = gD0floor;
    k.c = c; k.eps = clamp_eps; k.lo = clamp_lo;
    return k;
}

inline ConsConsts cons_consts(double dt, double a, double c, double gmax, double alpha, double rW, double W0, double K1, double K2,
                              double Dw, double Do, double A01, double H01, double E1)
{
    ConsConsts k{};
    k.dt = dt; k.dt2 = dt / 2.0; k.dt6 = dt / 6.0;                    // R/:171
    k.a = a; k.cgmax = c * gmax; k.gmax = gmax; k.alpha = alpha; k.rW = rW; k.K1 = K1; k.K2 = K2;
    k.K2W0 = K2 * W0; k.Dw = Dw; k.Do = Do;
    k.A01 = A01; k.A01H01 = A01 * H01; k.EA1 = E1 * A01;              // R/:2452
    return k;
}

// launch(kernel, args, grid, block, dynamic shared bytes); order and arguments as the reference's loop body R/:3752-4183:
// vegetation Dornic -> means and cover of DRho -> perception field -> grazer Dornic -> four RK4 stages (clamp in the last)
template <class Launch>
inline void cons_step2_enqueue(Launch &&launch, const Step2Consts &k, Step2Buffers &b, unsigned long long seed,
                               unsigned int trial, unsigned int step)
{
    const int L = k.L;
    const long long n = (long long)L * L;
    const int tiles = (L + kConsThreads - 1) / kConsThreads;
    const unsigned grid = (unsigned)((long long)tiles * L);
    const bool replay = b.replay_veg != nullptr;
    GrazerArgs V{};
    V.L = L; V.dt = k.dt; V.dx = k.dx; V.D = k.D; V.v = k.v; V.M = k.M; V.sigma = k.sigma; V.D0floor = k.D0floor;
    grazer_derive(V);
    V.prev = b.state; V.gamma = b.gamma; V.out = b.star; V.veff = b.veff;
    V.seed = seed; V.trial = trial; V.step = step; V.replay_u = b.replay_veg; V.stride = b.stride; V.overflow = b.overflow;
    V.veg = k.veg; V.dyn = nullptr;
    if (replay) launch(grazer_dornic_kernel<true, true>, V, grid, kConsThreads, 0);
    else        launch(grazer_dornic_kernel<false, true>, V, grid, kConsThreads, 0);
    if (b.strict_means) {
        SeqSumArgs S{n, {b.state, b.state + n, nullptr}, 2, b.part, kConsReduceBlocks};
        launch(cons_seqsum_kernel, S, 1, 32, 0);
    } else {
        ReduceArgs R{n, b.state, b.state + n, b.part};
        launch(cons_reduce_kernel, R, kConsReduceBlocks, kConsReduceThreads, sizeof(double) * 3 * kConsReduceThreads);
    }
    FinalizeArgs F{L, kConsReduceBlocks, k.n_central, k.fr, b.part, b.dyn, k.fpe};
    launch(cons_finalize_kernel, F, 1, 32, sizeof(double) * 96);
    GammaArgs G{};
    G.L = L; G.eff = k.n_central; G.prey = b.state; G.offsets = b.central; G.gamma = b.gamma; G.dyn = b.dyn;
    launch(cons_gamma_kernel, G, grid, kConsThreads, sizeof(int) * 2 * k.n_central);
    GrazerArgs Z = V;
    Z.prev = b.state + n; Z.out = b.star + n; Z.replay_u = b.replay_graz; Z.dyn = b.dyn;
    Z.seed = seed ^ kGrazerSeedSalt;
    if (replay) launch(grazer_dornic_kernel<true, false>, Z, grid, kConsThreads, 0);
    else        launch(grazer_dornic_kernel<false, false>, Z, grid, kConsThreads, 0);
    // Rietkerk_FD clamps in the RK4 commit (R/:4176); Rietkerk_FPE moves the grazer first and clamps after (F/:4308, :4458)
    if (!b.rk4_stages) {
        RkFusedArgs F4{};
        F4.L = L; F4.c = k.c; F4.base = b.state; F4.base_pg = b.star; F4.out = b.ya; F4.clamp = k.fpe ? 0 : 1; F4.eps = k.eps; F4.lo = k.lo;
        const int ft = (L + kFuseT - 1) / kFuseT;
        launch(cons_rk4_fused_kernel, F4, (unsigned)(ft * ft), kFuseThreads, sizeof(double) * 4 * kFuseS * kFuseS);
        double *t = b.state; b.state = b.ya; b.ya = t;
    } else {
        RkStageArgs A{};
        A.L = L; A.c = k.c; A.base = b.state; A.base_pg = b.star; A.acc = b.acc; A.out = b.state;
        A.clamp = k.fpe ? 0 : 1; A.eps = k.eps; A.lo = k.lo;
        A.stage = 1; A.Y = b.state; A.Ynext = b.ya; launch(cons_rk4_stage_kernel, A, grid, kConsThreads, 0);
        A.stage = 2; A.Y = b.ya; A.Ynext = b.yb; launch(cons_rk4_stage_kernel, A, grid, kConsThreads, 0);
        A.stage = 3; A.Y = b.yb; A.Ynext = b.ya; launch(cons_rk4_stage_kernel, A, grid, kConsThreads, 0);
        A.stage = 4; A.Y = b.ya; A.Ynext = nullptr; launch(cons_rk4_stage_kernel, A, grid, kConsThreads, 0);
    }
    if (!k.fpe) return;
    // advdiff_FKE_MultiSp (F/:4308) on the grazer field the RK4 left: three launches (tm_fke.cuh), scratch in the buffers the RK4 is
    // done with (yb: the three source masses and U1; acc: the moved field, U2, the displacements), then the commit + clamp
    FkeArgs M{};
    M.t = k.fke;
    M.rho = b.state + n; M.gamma = b.gamma; M.veff = b.veff; M.out = b.acc;
    M.sA = b.yb; M.sB = b.yb + n; M.sD = b.yb + 2 * n; M.U1 = b.yb + 3 * n; M.U2 = b.acc + n;
    M.disp = reinterpret_cast<int *>(b.acc + 2 * n); M.maxdisp = b.dyn->maxdisp;
    M.gD = b.fke_gD; M.gV = b.fke_gV;
    fke_move_enqueue(launch, M);
    FpeCommitArgs C{n, b.state, b.acc, k.eps, k.lo};
    launch(cons_fpe_commit_kernel, C, (unsigned)((n + kConsThreads - 1) / kConsThreads), kConsThreads, 0);
}
inline int step2_launches(const Step2Buffers &b, int fpe = 0) { return (b.rk4_stages ? 9 : 6) + (fpe ? 4 : 0); }

} // namespace tml
