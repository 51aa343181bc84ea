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
#include <type_traits>

#if defined(__CUDACC__)
#define TMC_D __device__ __forceinline__
#define TMC_G __global__
#define TMC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#include "tm_cpu_emu.h" // tests/emu only
#define TMC_D inline
#define TMC_G
#define TMC_SMEM(name) unsigned char *name = tm_emu_dyn_smem()
#endif

namespace tml {

constexpr int kConsThreads = 128;

struct GammaArgs {
    int L, eff;                  // lattice side; number of disc offsets summed (R/:2569)
    const double *prey;          // [L*L] Rho_t[s-1]
    const int *offsets;          // [eff][2] = {dx (rows), dy (columns)}
    double rho_inverse;          // 1 / max(<prey>, 1e-12)  (R/:2548)
    int zero;                    // <self> or <prey> below 1e-12: gamma = 0 everywhere (R/:2575-2578)
    double *gamma;               // [L*L] in/out
};

TMC_G void cons_gamma_kernel(const GammaArgs A)
{
    const int g = A.L, tid = threadIdx.x;
    const int tiles = (g + kConsThreads - 1) / kConsThreads;
    const int c_i = (int)(blockIdx.x / tiles), c_j = (int)(blockIdx.x % tiles) * kConsThreads + tid;
    TMC_SMEM(smem_raw);
    int *off = reinterpret_cast<int *>(smem_raw);
    for (int k = tid; k < 2 * A.eff; k += kConsThreads) off[k] = A.offsets[k];
    __syncthreads();
    if (c_j >= g) return;
    const int i = c_i * g + c_j;
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
    const double *prev;                       // DRho[s] [L*L]
    double *gamma;                            // [L*L] in/out (clamped at 1, R/:3915)
    double *out, *veff;                       // rho* [L*L]; advection vectors [L*L][2]
    unsigned long long seed; unsigned int trial, step;
    const double *replay_u; unsigned int stride;   // REPLAY: [L*L][stride]
    unsigned long long *overflow;
};

template <bool REPLAY>
TMC_G void grazer_dornic_kernel(const GrazerArgs A)
{
    const int g = A.L;
    const int tiles = (g + kConsThreads - 1) / kConsThreads;
    const int r = (int)(blockIdx.x / tiles), q = (int)(blockIdx.x % tiles) * kConsThreads + (int)threadIdx.x;
    if (q >= g) return;
    const int i = r * g + q;
    const int up = (r == 0 ? g - 1 : r - 1), dn = (r == g - 1 ? 0 : r + 1), lf = (q == 0 ? g - 1 : q - 1), rt = (q == g - 1 ? 0 : q + 1);
    const double dx2 = A.dx * A.dx, dx1 = 1 / A.dx, dx1_2 = 1 / (dx2), dxb2 = A.dx / 2.0;   // R/:3455
    const double sigma2_1 = 1 / (A.sigma * A.sigma);
    double gm = A.gamma[i];
    if (gm > 1) { gm = 1; A.gamma[i] = gm; }
    typename std::conditional<REPLAY, ReplaySrc, PhiloxSrc>::type src;
    if constexpr (REPLAY) src.begin(A.replay_u + (size_t)i * A.stride, A.stride);
    else src.begin((uint32_t)i, A.trial, A.step, A.seed);
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
};

struct RkStageArgs {
    int L, stage;                // stage 1..4
    ConsConsts c;
    const double *base;          // Rho_t   [4][L*L]  (P, G, W, O)
    const double *Y;             // stage state (== base for stage 1)
    double *Ynext;               // next stage state (stages 1-3)
    double *acc;                 // K1 (+ 2 K2 (+ 2 K3)) [4][L*L]
    double *out;                 // stage 4: the committed state (may alias base)
};

// f_2Dor_2Sp R/:2460-2468 at one site, the reference's operation order
TMC_D void cons_rhs(const ConsConsts &c, double P, double G, double W, double O, double Wu, double Wd, double Wl, double Wr,
                    double Ou, double Od, double Ol, double Or_, double f[4])
{
    f[0] = c.cgmax * W * P / (W + c.K1) - (c.A01 * P / (1 + c.A01H01 * P)) * G;
    f[1] = (c.EA1 * P / (1 + c.A01H01 * P)) * G;
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
    double f[4];
    cons_rhs(A.c, Y[i], Y[n + i], W[i], O[i], W[up], W[dn], W[lf], W[rt], O[up], O[dn], O[lf], O[rt], f);
    const double h = A.stage == 3 ? A.c.dt : A.c.dt2;
    for (int s = 0; s < 4; s++) {
        const int k = s * n + i;
        if (A.stage == 1) { A.acc[k] = f[s]; A.Ynext[k] = A.base[k] + (h) * f[s]; }
        else if (A.stage < 4) { A.acc[k] = A.acc[k] + 2.0 * f[s]; A.Ynext[k] = A.base[k] + (h) * f[s]; }
        else A.out[k] = A.base[k] + (A.c.dt6) * (A.acc[k] + f[s]);    // Rho_t += dt6 (K1 + 2 K2 + 2 K3 + K4), R/:2506
    }
}

} // namespace tml
