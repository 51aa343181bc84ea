// tm_dpdet.cuh -- the noise-off DP path (second half): percNoDiff_2D P/:549-662 = classical RK4 over
// fP_2D P/:516-529, drho/dt = a rho - b rho^2 + D lap4(rho) with the fourth-order nine-point-cross Laplacian
// (radius 2, determine_neighbours_R2 P/:394-414: [0][j] = rows p-2..p+2 of column q, [1][j] = columns q-2..q+2 of row p,
// periodic).  One thread per cell, one launch per RK stage, the reference's operation order (the TU is built with
// --fmad=false): bit-identical to the oracle.  HBM-bound: 3-5 field passes per stage.  P/ = Percolation_FD/stochasticSPDP.cpp.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TMN_G __global__
#else
#include "tm_cpu_emu.h" // tests/emu only
#define TMN_G
#endif

namespace tml {

constexpr int kDpDetThreads = 128;

struct DpDetArgs {
    int L, stage;                // stage 1..4
    double a, b, D, dx, dt;
    const double *base;          // Rho_dt  [L*L]
    const double *Y;             // stage state (== base for stage 1)
    double *Ynext;               // Rho_M of the next stage (stages 1-3)
    double *acc;                 // K1 (+ 2 K2 (+ 2 K3))
    double *out;                 // stage 4: Rho_dt after the step (may alias base)
};

TMN_G void dp_det_stage_kernel(const DpDetArgs A)
{
    const int g = A.L;
    const int tiles = (g + kDpDetThreads - 1) / kDpDetThreads;
    const int p = (int)(blockIdx.x / tiles), q = (int)(blockIdx.x % tiles) * kDpDetThreads + (int)threadIdx.x;
    if (q >= g) return;
    auto w = [g](int v) { return ((v % g) + g) % g; };                  // P/:406-407
    const double *R = A.Y;
    const int i = p * g + q;
    const double v0 = R[w(p - 2) * g + q], v1 = R[w(p - 1) * g + q], v3 = R[w(p + 1) * g + q], v4 = R[w(p + 2) * g + q];
    const double h0 = R[p * g + w(q - 2)], h1 = R[p * g + w(q - 1)], h3 = R[p * g + w(q + 1)], h4 = R[p * g + w(q + 2)];
    const double c = R[i];                                              // n[i][0][2]
    const double f = A.a * c - A.b * c * c - A.D * (1 / (12.0 * A.dx * A.dx)) * (1 * (v0 + v4 + h0 + h4) - 16 * (v1 + v3 + h1 + h3) + 60 * c);   // P/:522-524
    if (A.stage == 1) { A.acc[i] = f; A.Ynext[i] = A.base[i] + (A.dt / 2.0) * f; }                  // P/:592-593
    else if (A.stage == 2) { A.acc[i] = A.acc[i] + 2 * f; A.Ynext[i] = A.base[i] + (A.dt / 2.0) * f; }
    else if (A.stage == 3) { A.acc[i] = A.acc[i] + 2 * f; A.Ynext[i] = A.base[i] + (A.dt) * f; }
    else A.out[i] = A.base[i] + (A.dt / 6.0) * (A.acc[i] + f);                                       // P/:620
}

// one RK4 step as four launches; launch(kernel, args, grid, block, dynamic shared bytes)
template <class Launch>
inline void dp_det_step_enqueue(Launch &&launch, int L, double a, double b, double D, double dt, double dx, double *state,
                                double *ya, double *yb, double *acc)
{
    DpDetArgs A{};
    A.L = L; A.a = a; A.b = b; A.D = D; A.dx = dx; A.dt = dt; A.base = state; A.acc = acc; A.out = state;
    const unsigned grid = (unsigned)((long long)((L + kDpDetThreads - 1) / kDpDetThreads) * L);
    A.stage = 1; A.Y = state; A.Ynext = ya; launch(dp_det_stage_kernel, A, grid, kDpDetThreads, 0);
    A.stage = 2; A.Y = ya; A.Ynext = yb; launch(dp_det_stage_kernel, A, grid, kDpDetThreads, 0);
    A.stage = 3; A.Y = yb; A.Ynext = ya; launch(dp_det_stage_kernel, A, grid, kDpDetThreads, 0);
    A.stage = 4; A.Y = ya; A.Ynext = nullptr; launch(dp_det_stage_kernel, A, grid, kDpDetThreads, 0);
}

} // namespace tml
