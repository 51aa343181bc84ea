// tm_fke.cuh -- sm_100a kernels of the FKE movement step.
//
// What it replaces: updateMovement, simulations/Rietkerk_FPE/kernel_advdiff_FKE2D.cu:92-180 (the reference's only
// GPU kernel: one thread per SOURCE site scattering FP64 atomics into four rectangles, O(L^2 l^2) atomics, sum
// order non-deterministic) and its CPU twin advdiff_FKE_MultiSp, F/:3614-3695.
//
// Per consumer species, a live source i = (c_i, c_j) (rho >= epsilon) spreads its mass over
//   A  gamma^2          rho  gD|p - c_i|  gD|q - c_j|     p in c_i +- bD,               q in c_j +- bD
//   B  (gamma-gamma^2)  rho  gD|p - c_i|  gV|q - mu_y|    p in c_i +- bD,               q in [mu_y - clo, mu_y + chi]
//   C  (gamma-gamma^2)  rho  gV|p - mu_x| gD|q - c_j|     p in [mu_x - rlo, mu_x + rhi], q in c_j +- bD
//   D  (1-gamma)^2      rho  gV|p - mu_x| gV|q - mu_y|    p in [mu_x - rlo, mu_x + rhi], q in [mu_y - clo, mu_y + chi]
// with mu = round(c + v_eff * prefactor) mod L, all indices periodic.  Every term is "move a point mass by an integer
// displacement e = mu - c (zero for the undisplaced axis), then convolve with a fixed separable stencil", so the
// scatter is restated as a gather:
//   fke_prepare_kernel   sA, sB, sD and the packed displacement (e_x, e_y) of every site; max |e| per axis
//   fke_cols_kernel      T_A = sA, T_B / T_C / T_D = masses displaced along columns / rows / both, gathered from the
//                        (2 R_x + 1)(2 R_y + 1) candidate sources around each target (R = max |e|, usually 0..2),
//                        then the column convolutions U1 = gD *c T_A + gV *c T_B,  U2 = gD *c T_C + gV *c T_D
//   fke_rows_kernel      out = gD *r U1 + gV *r U2
// O(L^2 (bD + bV)) work instead of O(L^2 (bD + bV)^2), no atomics on the data path, a fixed summation order (results
// are bit-reproducible run to run) and all terms non-negative, so the result agrees with the reference's scatter order
// to a few ulp.  HBM-bound: 32 B read + 8 B written per site algorithmically; the three passes move ~130 B.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define TMF_D __device__ __forceinline__
#define TMF_HD __host__ __device__ inline
#define TMF_G static __global__   // static: tm_fke.cu and tm_consumer.cu (the FPE flavour of the step) each compile their own copy
#define TMF_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#include "tm_cpu_emu.h" // tests/emu only: runs the same source on host threads
#define TMF_D inline
#define TMF_HD inline
#define TMF_G
#define TMF_SMEM(name) unsigned char *name = tm_emu_dyn_smem()
#endif

namespace tml {

struct FkeTerms {
    int L, bD;                 // lattice side; sig_D_ScaledBounds[s]
    int rlo, rhi, clo, chi;    // advection window: rows [mu_x - rlo, mu_x + rhi], columns [mu_y - clo, mu_y + chi]
    double pref, eps;          // mu_vel_prefactor[s] (F/:249), epsilon (1e-10)
};

struct FkeArgs {
    FkeTerms t;
    const double *rho, *gamma, *veff;   // [L*L], [L*L], [L*L][2]
    double *sA, *sB, *sD;               // [L*L] source masses of terms A, B (= C), D; 0 for dead sites
    int *disp;                          // [L*L] (e_x & 0xFFFF) | (e_y << 16)
    int *maxdisp;                       // [2] max |e_x|, max |e_y| over live sites (zeroed by the caller)
    double *U1, *U2;                    // [L*L] column-convolved intermediates
    double *out;                        // [L*L]
    const double *gD, *gV;              // [bD + 1], [max(rlo, rhi, clo, chi) + 1]
};

constexpr int kFkeThreads = 128;        // columns per CTA in the two convolution passes
constexpr int kFkeMaxBound = 512;       // largest stencil half-width supported (shared-memory tile)
constexpr int kFkeMaxDisp = 64;         // largest |displacement| (cells per step, either axis) the gather-form column pass accepts: it
                                        // visits (2 R_x + 1)(2 R_y + 1) candidate sources per target, fine for the reference's fractions
                                        // of a cell, minutes per step for an outlier of hundreds of cells.  Beyond it the pass REFUSES:
                                        // the moved field comes back as NaN (the runners' non-finite guards then stop the run, as the
                                        // reference's own numerical-failure exit does) instead of occupying the GPU.

TMF_D int fke_pack(int ex, int ey) { return (ex & 0xFFFF) | (int)((unsigned)ey << 16); }
TMF_D int fke_mod(int a, int g) { int r = a % g; return r < 0 ? r + g : r; }
TMF_HD int fke_max3(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }

// F/:3622-3637: skip rule, the three source masses, and the integer displacement of the advected copies
TMF_G void fke_prepare_kernel(const FkeArgs A)
{
    const int g = A.t.L;
    int mx = 0, my = 0;
    // a CTA walks whole rows (no division to recover (row, column) from a flat index)
    for (int c_i = blockIdx.x; c_i < g; c_i += gridDim.x) {
        for (int c_j = threadIdx.x; c_j < g; c_j += blockDim.x) {
            const int i = c_i * g + c_j;
            const double r = A.rho[i];
            double a = 0.0, b = 0.0, d = 0.0;
            int ex = 0, ey = 0;
            if (!(r < A.t.eps)) {
                const double gm = A.gamma[i], gm2 = gm * gm;
                a = gm2 * r; b = (gm - gm2) * r; d = (1 - gm) * (1 - gm) * r;
                // (int)(round(c + v * prefactor) + g) % g in the reference; a true modulo here (the reference indexes
                // out of bounds when the left side is negative, i.e. for |v * prefactor| > L)
                // a non-finite or absurd shift (|v * prefactor| >= 1e9 cells: the reference's int conversion is undefined there)
                // moves nothing instead of indexing from garbage
                double sx = A.veff[2 * i] * A.t.pref, sy = A.veff[2 * i + 1] * A.t.pref;
                if (!(sx > -1.0e9 && sx < 1.0e9)) sx = 0.0;
                if (!(sy > -1.0e9 && sy < 1.0e9)) sy = 0.0;
                const int mu_x = fke_mod((int)(round(c_i + sx) + g), g);
                const int mu_y = fke_mod((int)(round(c_j + sy) + g), g);
                ex = mu_x - c_i; ey = mu_y - c_j;                    // representative in [-(g-1)/2, g/2]
                if (ex > g / 2) ex -= g; else if (ex < -((g - 1) / 2)) ex += g;
                if (ey > g / 2) ey -= g; else if (ey < -((g - 1) / 2)) ey += g;
                const int ax = ex < 0 ? -ex : ex, ay = ey < 0 ? -ey : ey;
                mx = ax > mx ? ax : mx; my = ay > my ? ay : my;
            }
            A.sA[i] = a; A.sB[i] = b; A.sD[i] = d;
            A.disp[i] = fke_pack(ex, ey);
        }
    }
    if (mx > 0) atomicMax(&A.maxdisp[0], mx);
    if (my > 0) atomicMax(&A.maxdisp[1], my);
}

// one CTA = one lattice row x, kFkeThreads output columns; shared: T_A..T_D over the columns +- Hc, the two stencils
TMF_HD size_t fke_cols_smem_doubles(const FkeTerms &t)
{
    const int Hc = fke_max3(t.bD, t.clo, t.chi);
    const int nV = fke_max3(fke_max3(t.rlo, t.rhi, t.clo), t.chi, 0) + 1;
    return (size_t)4 * (kFkeThreads + 2 * Hc) + (t.bD + 1) + nV;
}

TMF_G void fke_cols_kernel(const FkeArgs A)
{
    const FkeTerms &t = A.t;
    const int g = t.L, tid = threadIdx.x;
    const int tiles = (g + kFkeThreads - 1) / kFkeThreads;
    const int x = (int)(blockIdx.x / tiles), y0 = (int)(blockIdx.x % tiles) * kFkeThreads;
    const int Hc = fke_max3(t.bD, t.clo, t.chi);
    const int W = kFkeThreads + 2 * Hc;
    const int nV = fke_max3(fke_max3(t.rlo, t.rhi, t.clo), t.chi, 0) + 1;
    TMF_SMEM(smem_raw);
    double *TA = reinterpret_cast<double *>(smem_raw), *TB = TA + W, *TC = TB + W, *TD = TC + W;
    double *gD = TD + W, *gV = gD + (t.bD + 1);
    for (int k = tid; k <= t.bD; k += kFkeThreads) gD[k] = A.gD[k];
    for (int k = tid; k < nV; k += kFkeThreads) gV[k] = A.gV[k];
    const int Rx = A.maxdisp[0], Ry = A.maxdisp[1];
    if (Rx > kFkeMaxDisp || Ry > kFkeMaxDisp) {            // uniform over the launch (see kFkeMaxDisp): poison, do not gather
        const int yq = y0 + tid;
        if (yq < g) { A.U1[x * g + yq] = nan(""); A.U2[x * g + yq] = nan(""); }
        return;
    }
    // Index arithmetic: 32-bit throughout (L <= 32767, so L*L < 2^30) and periodic indices walked with a conditional
    // wrap -- an integer division and a 64-bit address per candidate source made this pass instruction-bound
    // (1100 warp instructions per 32 sites, 0.85 ms at L = 4096).  Candidates are read unconditionally and selected,
    // so the gather has no branches.
    const int rowx = x * g;
    const int xs0 = fke_mod(x + Rx, g);                    // first candidate row: x - ex for ex = -Rx
    const int ybase = fke_mod(y0 - Hc, g);
    const int *__restrict__ disp = A.disp;
    const double *__restrict__ sB = A.sB, *__restrict__ sD = A.sD;
    for (int k = tid; k < W; k += kFkeThreads) {
        int y = ybase + k; if (y >= g) y %= g;
        int ys0 = y + Ry; if (ys0 >= g) ys0 -= g;          // first candidate column: y - ey for ey = -Ry  (Ry <= g/2)
        double b = 0.0, c = 0.0, d = 0.0;
        // sources whose displaced copy lands on (x, y): (x - ex, y - ey) with disp == (ex, ey); ex = 0 for T_B, ey = 0 for T_C.
        // The undisplaced axis is NOT tested: term B uses only e_y of the source, term C only e_x (F/:3656, :3672).
        for (int ey = -Ry, ys = ys0; ey <= Ry; ey++) {
            const int s = rowx + ys;
            b += ((disp[s] >> 16) == ey) ? sB[s] : 0.0;
            ys = ys == 0 ? g - 1 : ys - 1;
        }
        for (int ex = -Rx, xs = xs0; ex <= Rx; ex++) {
            const int rs = xs * g;
            c += ((int)(short)(disp[rs + y] & 0xFFFF) == ex) ? sB[rs + y] : 0.0;
            for (int ey = -Ry, ys = ys0; ey <= Ry; ey++) {
                d += (disp[rs + ys] == fke_pack(ex, ey)) ? sD[rs + ys] : 0.0;
                ys = ys == 0 ? g - 1 : ys - 1;
            }
            xs = xs == 0 ? g - 1 : xs - 1;
        }
        TA[k] = A.sA[rowx + y]; TB[k] = b; TC[k] = c; TD[k] = d;
    }
    __syncthreads();
    const int y = y0 + tid;
    if (y < g) {
        const int k = tid + Hc;
        double u1 = 0.0, u2 = 0.0;
        for (int d = -t.bD; d <= t.bD; d++) {            // target q = c_j + d  <=>  source column k - d
            const double w = gD[d < 0 ? -d : d];
            u1 += w * TA[k - d]; u2 += w * TC[k - d];
        }
        for (int d = -t.clo; d <= t.chi; d++) {          // target q = mu_y + d
            const double w = gV[d < 0 ? -d : d];
            u1 += w * TB[k - d]; u2 += w * TD[k - d];
        }
        A.U1[rowx + y] = u1; A.U2[rowx + y] = u2;
    }
}

// out(x, y) = sum_d gD|d| U1(x - d, y) + sum_{d = -rlo..rhi} gV|d| U2(x - d, y); a CTA covers kFkeThreads columns of one row
TMF_G void fke_rows_kernel(const FkeArgs A)
{
    const FkeTerms &t = A.t;
    const int g = t.L, tid = threadIdx.x;
    const int tiles = (g + kFkeThreads - 1) / kFkeThreads;
    const int x = (int)(blockIdx.x / tiles), y = (int)(blockIdx.x % tiles) * kFkeThreads + tid;
    const int nV = fke_max3(t.rlo, t.rhi, 0) + 1;
    TMF_SMEM(smem_raw);
    double *gD = reinterpret_cast<double *>(smem_raw), *gV = gD + (t.bD + 1);
    for (int k = tid; k <= t.bD; k += kFkeThreads) gD[k] = A.gD[k];
    for (int k = tid; k < nV; k += kFkeThreads) gV[k] = A.gV[k];
    __syncthreads();
    if (y >= g) return;
    double acc = 0.0;
    for (int d = -t.bD, xs = fke_mod(x + t.bD, g); d <= t.bD; d++) {       // source row x - d, walked with a conditional wrap
        acc += gD[d < 0 ? -d : d] * A.U1[xs * g + y];
        xs = xs == 0 ? g - 1 : xs - 1;
    }
    for (int d = -t.rlo, xs = fke_mod(x + t.rlo, g); d <= t.rhi; d++) {
        acc += gV[d < 0 ? -d : d] * A.U2[xs * g + y];
        xs = xs == 0 ? g - 1 : xs - 1;
    }
    A.out[x * g + y] = acc;
}

// the three launches of one species' movement step (host side; shared by tm_fke.cu, the FPE flavour of the consumer stepper and
// tests/emu): launch(kernel, args, grid, block, dynamic shared bytes).  A.maxdisp must hold zeros on entry.
template <class Launch>
inline void fke_move_enqueue(Launch &&launch, const FkeArgs &A)
{
    const int L = A.t.L;
    const int tiles = (L + kFkeThreads - 1) / kFkeThreads;
    launch(fke_prepare_kernel, A, (unsigned)L, 256, 0);   // one CTA per lattice row
    launch(fke_cols_kernel, A, (unsigned)((long long)tiles * L), kFkeThreads, fke_cols_smem_doubles(A.t) * sizeof(double));
    launch(fke_rows_kernel, A, (unsigned)((long long)tiles * L), kFkeThreads,
           (size_t)(A.t.bD + 1 + fke_max3(A.t.rlo, A.t.rhi, 0) + 1) * sizeof(double));
}

} // namespace tml
