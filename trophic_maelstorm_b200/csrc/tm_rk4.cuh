// tm_rk4.cuh -- riet_rk4_kernel: the RK4 fronts of the Rietkerk step (R/:2389-2440 with f_2Dor R/:2359-2386, clamp and
// commit R/:4150-4181) for the production pair "riet_dornic_kernel, then this kernel".  sm_100a only.
//
// Same marching idea as riet_step_kernel (tm_kernels.cuh): a CTA owns a column strip and a run of rows, input rows stream
// into a shared-memory ring through TMA bulk copies, and at row step j four fronts advance at once
//     row j-1: K1      row j-2: K2      row j-3: K3      row j-4: K4, clamp, commit.
// What is different (round 2; the round-1 kernel issued 272 warp instructions per 32 site-updates, 129 of them FP64, and
// sat at 58 % of the FP64 pipe behind a per-row barrier with thread 0 issuing all copies after it):
//   * a thread owns TWO ADJACENT columns.  Of the four horizontal neighbours a front needs per field, two are the thread's
//     own registers; the other two come from the exchange rows as one 128-bit {W,O} load each: half the shared-memory
//     instructions and wavefronts per site, and two independent dependency chains per thread for the FP64 pipe.
//   * everything a thread made itself stays in registers: the centre and upper-neighbour values of every stage, the P stage
//     values, the K sums, the stage-0 centre/upper rows.  Shared memory carries only what crosses threads (two-slot exchange
//     rows of {W,O} pairs per stage) and the streamed input rows.
//   * the bulk copies of a row are issued by three different warps (lane 0 of warps 0, 1, 2 takes rho*, W, O), each with its own
//     arrive.expect_tx on the row's mbarrier (count 3): no warp is the straggler of the step barrier.
//   * strips are 2 x 32 x kRkWarps columns wide (256 with 4 warps): 3 % of the columns are recomputed halo instead of 5 %.
// Per step and thread (2 site-updates): 20 shared loads, 6 shared stores, 3 128-bit global stores, one CTA barrier.
// HBM traffic: rho*, W, O read once, P, W, O written once = 48 B per site-update; halo columns/rows are L2 hits.
#pragma once
#include "tm_kernels.cuh"

namespace tml {

#ifndef TM_RK4_WARPS
#define TM_RK4_WARPS 4
#endif
#ifndef TM_RK4_MINBLOCKS
#define TM_RK4_MINBLOCKS 3
#endif
#ifndef TM_RK4_SLOTS
#define TM_RK4_SLOTS 8      // ring slots (power of two): rows j-4 .. j live, rows j+1 .. j+kRkLead in flight
#endif
constexpr int kRkWarps = TM_RK4_WARPS;
constexpr int kRkThreads = 32 * kRkWarps;
constexpr int kRkSW = 2 * kRkThreads;           // columns a strip touches (halo included)
constexpr int kRkTXMax = kRkSW - 2 * kHalo;     // output columns per strip
#ifndef TM_RK4_CARRY0U
#define TM_RK4_CARRY0U 1    // stage-0 upper rows (W, O of row j-2) carried in registers (1) or re-read from the input ring (0)
#endif
#ifndef TM_RK4_UNROLL
#define TM_RK4_UNROLL 3     // row steps per loop trip: the register windows (own centre / upper values: period 3) rotate by renaming
#endif
constexpr int kRkUnroll = TM_RK4_UNROLL;
constexpr int kRkNS = TM_RK4_SLOTS;
constexpr int kRkLead = kRkNS - 5;
constexpr int kRkRW = kRkSW + 4;                // ring row: 2 pad doubles each side (TMA destination 16-B aligned)
constexpr int kRkXW = kRkSW + 2;                // exchange row: one pad {W,O} each side
static_assert(kRkWarps >= 3, "three warps issue the three bulk copies of a row");
static_assert((kRkNS & (kRkNS - 1)) == 0 && kRkLead >= 1, "ring slots: a power of two, at least 6");

TM_HD size_t riet_rk4_smem_bytes()
{
    return sizeof(double) * (size_t)kRkNS * 3 * kRkRW + 16 * (size_t)3 * 2 * kRkXW + sizeof(uint64_t) * kRkNS + 128;
}

struct alignas(16) D2 { double x, y; };         // a thread's two columns of one row of one field

struct Rk4Args {
    Geom g;                        // g.ns / g.TX are NOT used here: ns2 / TX2 below
    RietConsts c;
    const double *in;              // state: fields W (1) and O (2) are read
    const double *pstar;           // rho* of riet_dornic_kernel, padded layout [n_trials][rows_alloc][Lp]
    double *out;
    const double *a_trial;
    unsigned long long *counters;  // [1] pipeline timeouts
    int ns2, TX2;                  // strips per row and output columns per strip (even)
};

// strips of the two-column kernel: as few as possible, evened out, even width
static inline void choose_strips2(int L, int *ns_out, int *tx_out)
{
    int ns = (L + kRkTXMax - 1) / kRkTXMax;
    if (ns < 1) ns = 1;
    int tx = (((L + ns - 1) / ns) + 1) & ~1;
    if (tx > kRkTXMax) tx = kRkTXMax;
    *ns_out = ns; *tx_out = tx;
}

#if defined(__CUDACC__)
TM_D void mbar_arrive_expect_tx(uint64_t *b, uint32_t bytes) { mbar_expect_tx(b, bytes); }
TM_D D2 lds2(const double *p) { return *reinterpret_cast<const D2 *>(p); }
#else
inline D2 lds2(const double *p) { D2 v; v.x = p[0]; v.y = p[1]; return v; }
#endif

TM_HD int jlo0(int ya) { return ya - kHalo; }
TM_HD int imin(int a, int b) { return a < b ? a : b; }

template <bool STRICT>
#if defined(__CUDACC__)
__global__ void __launch_bounds__(kRkThreads, TM_RK4_MINBLOCKS)
#else
void
#endif
riet_rk4_kernel(const Rk4Args A)
{
    const Geom &g = A.g;
    const RietConsts &c = A.c;
    const int tid = threadIdx.x;
    constexpr int RW = kRkRW, XW = kRkXW, NS = kRkNS;
    TM_DYN_SMEM(smem_raw);
    double *ring = reinterpret_cast<double *>(smem_raw);            // [NS][3][RW]: rho*, W, O
    WO *xch = reinterpret_cast<WO *>(ring + NS * 3 * RW);           // [3 stages][2][XW]
    uint64_t *mbar = reinterpret_cast<uint64_t *>(xch + 3 * 2 * XW);

    if (tid == 0) {
        for (int s = 0; s < NS; s++) mbar_init(&mbar[s], 3);
        mbar_init_fence();
    }
    for (int i = tid; i < NS * 3 * RW + 2 * 3 * 2 * XW; i += kRkThreads) ring[i] = 0.0;  // ring and exchange rows start finite
    __syncthreads();

    const long long total = (long long)g.n_trials * A.ns2 * g.nrows;
    const long long w0 = total * blockIdx.x / gridDim.x;
    const long long w1 = total * (blockIdx.x + 1) / gridDim.x;
    const int c0 = 2 * tid;                                         // first of this thread's two strip columns
    const int eps_hi = hi32(c.eps);
    const double hP2 = STRICT ? c.dt2 : c.dt2cg, hP1 = STRICT ? c.dt : c.dtcg, hP6 = STRICT ? c.dt6 : c.dt6cg;
    const int warp = tid >> 5;
    const bool issuer = (tid & 31) == 0 && warp < 3;                // lane 0 of warps 0..2: field `warp` of every row
    // this thread's places in the rings: its pair in an input row; in an exchange row (one stage, one slot) the {W,O} of its
    // second column at xb[0] and of its first column at xa[0], where a row is laid out as
    //   [pad][second columns of threads 0..T-1][first columns of threads 0..T-1][pad]
    // so that both stores and both neighbour loads (xb[-1]: left thread's second column, xa[1]: right thread's first column)
    // are unit-stride 128-bit accesses across the warp: no bank conflicts (the interleaved layout had 2-way conflicts on all four)
    const double *const ringc = ring + 2 + c0;
    WO *const xb = xch + 1 + tid;
    WO *const xa = xch + 1 + kRkThreads + tid;

    unsigned int n = 0;                                             // rows consumed so far by this CTA (slot and parity counter)
    bool timeout = false;

    for (long long w = w0; w < w1;) {
        const int ts = (int)(w / g.nrows);
        const int trial = ts / A.ns2, strip = ts % A.ns2;
        const int ya = (int)(w % g.nrows);
        const int yb = (int)((w1 - w < (long long)(g.nrows - ya)) ? ya + (w1 - w) : g.nrows);
        w += yb - ya;

        const int x0 = strip * A.TX2;                               // first output column; the strip starts 4 columns to its left
        const int gx = x0 - kHalo + c0;                             // lattice column of the first of the two (ghost if < 0 or >= L)
        const bool in0 = (c0 >= kHalo) && (c0 < kHalo + A.TX2) && (gx < g.L);
        const bool in1 = (c0 + 1 >= kHalo) && (c0 + 1 < kHalo + A.TX2) && (gx + 1 < g.L);
        // ghost-column duty (the padded row repeats cells 0..3 after cell L-1 and cells L-4..L-1 before cell 0): warp-uniform flag
        const bool ghost_any = (in0 && (gx < kHalo || gx >= g.L - kHalo)) || (in1 && (gx + 1 < kHalo || gx + 1 >= g.L - kHalo));
#if defined(__CUDACC__)
        const bool warp_ghost = __any_sync(0xffffffffu, ghost_any);
#else
        const bool warp_ghost = true;
#endif
        int ncols = g.Lp - x0; if (ncols > kRkSW) ncols = kRkSW;    // columns of this strip that exist in the padded row
        const uint32_t row_bytes = (uint32_t)(ncols * sizeof(double));
        const double *fld = (warp == 0) ? A.pstar + (long long)trial * g.field_stride
                                        : A.in + (long long)trial * g.trial_stride + (long long)warp * g.field_stride;
        // commit pointer of the row front 4 is on (row jlo - 4 at the first step), walked down one row per step with the periodic
        // wrap; the ghost-column copies of a column sit at a fixed distance from it (+L for cells 0..3, -L for cells L-4..L-1)
        int r4 = jlo0(ya) - kHalo;
        if (g.wrap) { r4 %= g.L; if (r4 < 0) r4 += g.L; } else r4 += kHalo;
        double *po = A.out + (long long)trial * g.trial_stride + (long long)r4 * g.Lp + (gx + kHalo);
        const long long fs = g.field_stride;
        int gsh[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const bool inq = q ? in1 : in0;
            const int gq = gx + q;
            gsh[q] = !inq ? 0 : (gq < kHalo) ? g.L : (gq >= g.L - kHalo) ? -g.L : 0;
        }
        const double a_rain = A.a_trial[trial];
        const int jlo = ya - kHalo, jhi = yb + kHalo - 1;
        const int nit = (jhi - jlo + 1 + kRkUnroll - 1) / kRkUnroll * kRkUnroll;   // padded to the unroll: the dummy rows re-read row jhi and store nothing

        // lattice-local row j in [-4, nrows + 4) -> physical row
        auto prow = [&](int j) -> int {
            if (g.wrap) { if (j < 0) j += g.L; else if (j >= g.L) j -= g.L; return j; }
            return j + kHalo;
        };
        // Row requests: ring entry m takes lattice-local row jq (rows past jhi re-read row jhi: the dummy steps of the padded
        // loop).  The element offset of the row is carried and advanced by Lp per request (periodic wrap when the physical row
        // reaches L): nothing is rebuilt from kernel parameters on the issuing lanes' path.  Three bulk copies, one per issuing warp.
        unsigned int mq = n;
        int jq = jlo, rq = prow(jlo);
        long long offq = (long long)rq * g.Lp + x0;
        auto issue_next = [&]() {
            if (issuer) {
                const unsigned int slot = mq & (NS - 1);
                mbar_arrive_expect_tx(&mbar[slot], row_bytes);
                bulk_g2s(ring + (slot * 3 + warp) * RW + 2, fld + offq, row_bytes, &mbar[slot]);
            }
            mq++;
            if (jq < jhi) {
                jq++;
                if (g.wrap && rq + 1 == g.L) { rq = 0; offq -= (long long)(g.L - 1) * g.Lp; } else { rq++; offq += g.Lp; }
            }
        };
        for (int k = 0; k <= kRkLead && k < nit; k++) issue_next();

        // per-column state: [0], [1] = the thread's two columns
        double W0c[2] = {0, 0}, O0c[2] = {0, 0}, W0u[2] = {0, 0}, O0u[2] = {0, 0};       // stage 0, rows j-1 and j-2
        double n1w[2] = {0, 0}, n1o[2] = {0, 0}, u1w[2] = {0, 0}, u1o[2] = {0, 0};       // stage 1, rows j-2 and j-3
        double n2w[2] = {0, 0}, n2o[2] = {0, 0}, u2w[2] = {0, 0}, u2o[2] = {0, 0};       // stage 2, rows j-3 and j-4
        double n3w[2] = {0, 0}, n3o[2] = {0, 0}, u3w[2] = {0, 0}, u3o[2] = {0, 0};       // stage 3, rows j-4 and j-5
        double P1[2] = {0, 0}, P2[2] = {0, 0}, P3[2] = {0, 0};                           // P stage values of rows j-2, j-3, j-4
        double aP2[2] = {0, 0}, aW2[2] = {0, 0}, aO2[2] = {0, 0};                        // K sums of rows j-2, j-3, j-4
        double aP3[2] = {0, 0}, aW3[2] = {0, 0}, aO3[2] = {0, 0};
        double aP4[2] = {0, 0}, aW4[2] = {0, 0}, aO4[2] = {0, 0};

#pragma unroll 1
        for (int it0 = 0; it0 < nit; it0 += kRkUnroll) {
#pragma unroll
        for (int uu = 0; uu < kRkUnroll; uu++, n++) {
            const int it = it0 + uu;
            const int j = jlo + it;
            if (!timeout && !mbar_wait(&mbar[n & (NS - 1)], (n / NS) & 1u)) timeout = true;
            // ring rows j .. j-4 (this thread's pair) and the exchange rows read / written in this step
            constexpr unsigned kSlotB = 3 * RW * sizeof(double);
            const char *const rb = reinterpret_cast<const char *>(ringc);
            const double *r0 = reinterpret_cast<const double *>(rb + ((n) & (NS - 1)) * kSlotB);
            const double *r1 = reinterpret_cast<const double *>(rb + ((n - 1) & (NS - 1)) * kSlotB);
            const double *r2 = reinterpret_cast<const double *>(rb + ((n - 2) & (NS - 1)) * kSlotB);
            const double *r3 = reinterpret_cast<const double *>(rb + ((n - 3) & (NS - 1)) * kSlotB);
            const double *rr4 = reinterpret_cast<const double *>(rb + ((n - 4) & (NS - 1)) * kSlotB);
            const int xro = (n & 1u) * XW, xwo = ((n + 1) & 1u) * XW;   // exchange slot written during the previous step / read during the next

#if !TM_RK4_CARRY0U
            {   // stage-0 rows j-2 of this thread's pair re-read from the ring instead of carried (8 registers)
                const D2 wu = lds2(r2 + RW), ou = lds2(r2 + 2 * RW);
                W0u[0] = wu.x; W0u[1] = wu.y; O0u[0] = ou.x; O0u[1] = ou.y;
            }
#endif
            // ---- front 1: row j-1: K1 ------------------------------------------------------------------
            const D2 Ps1 = lds2(r1);
            const D2 Wd = lds2(r0 + RW), Od = lds2(r0 + 2 * RW);
            const double Wl = r1[RW - 1], Wr = r1[RW + 2], Ol = r1[2 * RW - 1], Or_ = r1[2 * RW + 2];
            const double Ps1v[2] = {Ps1.x, Ps1.y}, Wdv[2] = {Wd.x, Wd.y}, Odv[2] = {Od.x, Od.y};
            double W1[2], O1[2], P1n[2], k1p[2], k1w[2], k1o[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const F3 k = rhs<false, STRICT>(c, a_rain, Ps1v[q], W0c[q], O0c[q],
                                                W0u[q], Wdv[q], q ? W0c[0] : Wl, q ? Wr : W0c[1],
                                                O0u[q], Odv[q], q ? O0c[0] : Ol, q ? Or_ : O0c[1], 0, 0, 0, 0);
                P1n[q] = axpy<STRICT>(hP2, k.p, Ps1v[q]);
                W1[q] = axpy<STRICT>(c.dt2, k.w, W0c[q]);
                O1[q] = axpy<STRICT>(c.dt2, k.o, O0c[q]);
                k1p[q] = k.p; k1w[q] = k.w; k1o[q] = k.o;
            }
            xa[xwo] = WO{W1[0], O1[0]}; xb[xwo] = WO{W1[1], O1[1]};

            // ---- front 2: row j-2: K2 ------------------------------------------------------------------
            const D2 Ps2 = lds2(r2);
            const WO s1l = xb[xro - 1], s1r = xa[xro + 1];
            const double Ps2v[2] = {Ps2.x, Ps2.y};
            double W2[2], O2[2], P2n[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const F3 k = rhs<false, STRICT>(c, a_rain, P1[q], n1w[q], n1o[q],
                                                u1w[q], W1[q], q ? n1w[0] : s1l.w, q ? s1r.w : n1w[1],
                                                u1o[q], O1[q], q ? n1o[0] : s1l.o, q ? s1r.o : n1o[1], 0, 0, 0, 0);
                P2n[q] = axpy<STRICT>(hP2, k.p, Ps2v[q]);
                W2[q] = axpy<STRICT>(c.dt2, k.w, W0u[q]);
                O2[q] = axpy<STRICT>(c.dt2, k.o, O0u[q]);
                aP2[q] = axpy<STRICT>(2.0, k.p, aP2[q]); aW2[q] = axpy<STRICT>(2.0, k.w, aW2[q]); aO2[q] = axpy<STRICT>(2.0, k.o, aO2[q]);
            }
            xa[2 * XW + xwo] = WO{W2[0], O2[0]}; xb[2 * XW + xwo] = WO{W2[1], O2[1]};

            // ---- front 3: row j-3: K3 ------------------------------------------------------------------
            const D2 Ps3 = lds2(r3), Wb3 = lds2(r3 + RW), Ob3 = lds2(r3 + 2 * RW);
            const WO s2l = xb[2 * XW + xro - 1], s2r = xa[2 * XW + xro + 1];
            const double Ps3v[2] = {Ps3.x, Ps3.y}, Wb3v[2] = {Wb3.x, Wb3.y}, Ob3v[2] = {Ob3.x, Ob3.y};
            double W3[2], O3[2], P3n[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const F3 k = rhs<false, STRICT>(c, a_rain, P2[q], n2w[q], n2o[q],
                                                u2w[q], W2[q], q ? n2w[0] : s2l.w, q ? s2r.w : n2w[1],
                                                u2o[q], O2[q], q ? n2o[0] : s2l.o, q ? s2r.o : n2o[1], 0, 0, 0, 0);
                P3n[q] = axpy<STRICT>(hP1, k.p, Ps3v[q]);
                W3[q] = axpy<STRICT>(c.dt, k.w, Wb3v[q]);
                O3[q] = axpy<STRICT>(c.dt, k.o, Ob3v[q]);
                aP3[q] = axpy<STRICT>(2.0, k.p, aP3[q]); aW3[q] = axpy<STRICT>(2.0, k.w, aW3[q]); aO3[q] = axpy<STRICT>(2.0, k.o, aO3[q]);
            }
            xa[4 * XW + xwo] = WO{W3[0], O3[0]}; xb[4 * XW + xwo] = WO{W3[1], O3[1]};

            // ---- front 4: row j-4: K4, clamp, commit -------------------------------------------------------
            const D2 Ps4 = lds2(rr4), Wb4 = lds2(rr4 + RW), Ob4 = lds2(rr4 + 2 * RW);
            const WO s3l = xb[4 * XW + xro - 1], s3r = xa[4 * XW + xro + 1];
            const double Ps4v[2] = {Ps4.x, Ps4.y}, Wb4v[2] = {Wb4.x, Wb4.y}, Ob4v[2] = {Ob4.x, Ob4.y};
            double nP[2], nW[2], nO[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const F3 k = rhs<false, STRICT>(c, a_rain, P3[q], n3w[q], n3o[q],
                                                u3w[q], W3[q], q ? n3w[0] : s3l.w, q ? s3r.w : n3w[1],
                                                u3o[q], O3[q], q ? n3o[0] : s3l.o, q ? s3r.o : n3o[1], 0, 0, 0, 0);
                nP[q] = axpy<STRICT>(hP6, aP4[q] + k.p, Ps4v[q]);
                nW[q] = axpy<STRICT>(c.dt6, aW4[q] + k.w, Wb4v[q]);
                nO[q] = axpy<STRICT>(c.dt6, aO4[q] + k.o, Ob4v[q]);
            }
            {   // R/:4176 `if (v < eps && v >= lo) v = 0`, almost never true for a live cell: pre-tested on the high words of the six
                // values at once (v < eps implies hi(v) <= hi(eps); a set sign bit compares below)
                int m0 = hi32(nP[0]), m1 = hi32(nP[1]);
                m0 = imin(m0, hi32(nW[0])); m1 = imin(m1, hi32(nW[1])); m0 = imin(m0, hi32(nO[0])); m1 = imin(m1, hi32(nO[1]));
                if (imin(m0, m1) <= eps_hi) {
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        nP[q] = clamp_small(nP[q], c.eps, c.lo); nW[q] = clamp_small(nW[q], c.eps, c.lo); nO[q] = clamp_small(nO[q], c.eps, c.lo);
                    }
                }
            }
            if (j - 4 >= ya && j - 4 < yb) {
                if (in0 && in1) {
#if defined(__CUDACC__)
                    *reinterpret_cast<double2 *>(po) = make_double2(nP[0], nP[1]);
                    *reinterpret_cast<double2 *>(po + fs) = make_double2(nW[0], nW[1]);
                    *reinterpret_cast<double2 *>(po + 2 * fs) = make_double2(nO[0], nO[1]);
#else
                    po[0] = nP[0]; po[1] = nP[1]; po[fs] = nW[0]; po[fs + 1] = nW[1]; po[2 * fs] = nO[0]; po[2 * fs + 1] = nO[1];
#endif
                } else {
                    if (in0) { po[0] = nP[0]; po[fs] = nW[0]; po[2 * fs] = nO[0]; }
                    if (in1) { po[1] = nP[1]; po[fs + 1] = nW[1]; po[2 * fs + 1] = nO[1]; }
                }
                if (warp_ghost) {   // strip-edge warps only
#pragma unroll
                    for (int q = 0; q < 2; q++)
                        if (gsh[q] != 0) { double *pg = po + (q + gsh[q]); pg[0] = nP[q]; pg[fs] = nW[q]; pg[2 * fs] = nO[q]; }
                }
            }
            if (g.wrap && r4 + 1 == g.L) { r4 = 0; po -= (long long)(g.L - 1) * g.Lp; } else { r4 += 1; po += g.Lp; }
            // ---- shift the windows by one row ---------------------------------------------------------------
#pragma unroll
            for (int q = 0; q < 2; q++) {
#if TM_RK4_CARRY0U
                W0u[q] = W0c[q]; O0u[q] = O0c[q];
#endif
                W0c[q] = Wdv[q]; O0c[q] = Odv[q];
                u1w[q] = n1w[q]; u1o[q] = n1o[q]; n1w[q] = W1[q]; n1o[q] = O1[q];
                u2w[q] = n2w[q]; u2o[q] = n2o[q]; n2w[q] = W2[q]; n2o[q] = O2[q];
                u3w[q] = n3w[q]; u3o[q] = n3o[q]; n3w[q] = W3[q]; n3o[q] = O3[q];
                P3[q] = P3n[q]; P2[q] = P2n[q]; P1[q] = P1n[q];
                aP4[q] = aP3[q]; aW4[q] = aW3[q]; aO4[q] = aO3[q];
                aP3[q] = aP2[q]; aW3[q] = aW2[q]; aO3[q] = aO2[q];
                aP2[q] = k1p[q]; aW2[q] = k1w[q]; aO2[q] = k1o[q];
            }
            __syncthreads();
            // everyone is done with ring row j-4: its slot takes row j + kRkLead + 1
            if (it + kRkLead + 1 < nit) issue_next();
        }
        }
    }
    if (timeout && tid == 0) atomicAdd(&A.counters[1], 1ull);
}


// ------------------------------------------------------------------------------------------------------------------
// riet_dornic2_kernel: the vegetation Dornic sweep (R/:3752-3809) with TWO adjacent columns per thread.
// riet_dornic_kernel (one column per thread, tm_kernels.cuh) spends 39 % of its warp-cycles in fixed-latency dependency
// stalls (ncu, round 2): a cell is one serial chain Philox -> log -> sqrt -> sincos -> truncation -> sqrt with almost no
// parallelism inside it.  Two cells per thread are two independent chains the scheduler can interleave; the centre pair is
// one 128-bit load and one 128-bit store per row.  Same arguments, same results (riet_dornic_value per cell), same
// second walk for the cells that need the exact samplers.
// ------------------------------------------------------------------------------------------------------------------
#ifndef TM_DORNIC2_MINBLOCKS
#define TM_DORNIC2_MINBLOCKS 3
#endif
#ifndef TM_DORNIC2_THREADS
#define TM_DORNIC2_THREADS 256
#endif
constexpr int kDornic2Threads = TM_DORNIC2_THREADS;
constexpr int kDornic2Cols = 2 * kDornic2Threads;     // columns per CTA

template <bool STRICT>
#if defined(__CUDACC__)
__global__ void __launch_bounds__(kDornic2Threads, TM_DORNIC2_MINBLOCKS)
#else
void
#endif
riet_dornic2_kernel(const DornicArgs A)
{
    const Geom &g = A.g;
    const int tid = threadIdx.x;
    TM_DYN_SMEM(smem_raw);
    double *logtab = reinterpret_cast<double *>(smem_raw); // [128][2]
    double *sctab = logtab + 256;                          // [128][2]
    for (int i = tid; i < 256; i += kDornic2Threads) {
        if (!STRICT) logtab[i] = A.logtab[i];
        sctab[i] = sincos_table()[i];
    }
    __syncthreads();
    const int ncb = (g.L + kDornic2Cols - 1) / kDornic2Cols;
    const int nrow = A.row_hi - A.row_lo;
    const int nrun = (nrow + A.rows_per_cta - 1) / A.rows_per_cta;
    const long long items = (long long)g.n_trials * ncb * nrun;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int run = (int)(it % nrun);
        const int cb = (int)((it / nrun) % ncb);
        const int trial = (int)(it / ((long long)nrun * ncb));
        const int ra = A.row_lo + run * A.rows_per_cta;
        const int rb = (ra + A.rows_per_cta < A.row_hi) ? ra + A.rows_per_cta : A.row_hi;
        const int q0 = cb * kDornic2Cols + 2 * tid;      // first of the two columns (even)
        const bool v0 = q0 < g.L, v1 = q0 + 1 < g.L;
        // idle lanes (and the missing second column of an odd lattice) read a legal, aligned pair: the last whole pair of the row
        const int qc = v0 ? q0 : ((g.L - 1) & ~1);
        const bool edge_cta = (cb == 0) || (cb == ncb - 1);
        const uint32_t ptrial = A.trial0 + (uint32_t)trial;
        auto prow = [&](int r) -> int {
            if (g.wrap) { if (r < 0) r += g.L; if (r >= g.L) r -= g.L; return r; }
            return r + kHalo;
        };
        const long long col = (long long)kHalo + qc;
        const double *const inT = A.in + (long long)trial * g.trial_stride + col;
        const long long odelta = (A.pstar + (long long)trial * g.field_stride) - (A.in + (long long)trial * g.trial_stride);
        const double *pc = inT + (long long)prow(ra) * g.Lp, *p1 = inT + (long long)prow(ra + 1) * g.Lp;
        int pr2 = prow(ra + 2);
        const double *p2 = inT + (long long)pr2 * g.Lp;
        const long long wrap_back = (long long)g.L * g.Lp;
        D2 Pu = lds2(inT + (long long)prow(ra - 1) * g.Lp), Pc = lds2(pc), Pd = lds2(p1);
        double Pl = pc[-1], Pr = pc[2];                  // ghost columns cover q = 0 and q = L-1
        int grow = g.row0 + ra; if (grow < 0) grow += g.L; if (grow >= g.L) grow -= g.L;
        uint32_t cell = (uint32_t)grow * (uint32_t)g.L + (uint32_t)qc;
        int n_exact = 0;
        for (int row = ra; row < rb;) {
            int n_safe = rb - row;
            if (g.wrap && g.L - pr2 < n_safe) n_safe = g.L - pr2;
            if (g.L - grow < n_safe) n_safe = g.L - grow;
            const int row_end = row + n_safe;
#pragma unroll 2
            for (; row < row_end; row++) {
                D2 Pn = Pd; double Pln = Pl, Prn = Pr;
                if (row + 1 < rb) { Pn = lds2(p2); Pln = p1[-1]; Prn = p1[2]; }      // for the next iteration
                const double nb0 = STRICT ? Pu.x + Pd.x + Pl + Pc.y : (Pu.x + Pd.x) + (Pl + Pc.y);   // STRICT: the reference's order R/:3756
                const double nb1 = STRICT ? Pu.y + Pd.y + Pc.x + Pr : (Pu.y + Pd.y) + (Pc.x + Pr);
                bool ex0, ex1;
                double s0 = riet_dornic_value<STRICT>(A.dk, A.inv_lambda, A.mu_scale, Pc.x, nb0, v0, cell, ptrial, A.step, A.rk, logtab, sctab, ex0);
                double s1 = riet_dornic_value<STRICT>(A.dk, A.inv_lambda, A.mu_scale, Pc.y, nb1, v1, cell + 1u, ptrial, A.step, A.rk, logtab, sctab, ex1);
                if (ex0) { s0 = u2d(kExactMarker); n_exact++; }
                if (ex1) { s1 = u2d(kExactMarker); n_exact++; }
                double *pout = const_cast<double *>(pc) + odelta;
                if (v1) {
#if defined(__CUDACC__)
                    *reinterpret_cast<double2 *>(pout) = make_double2(s0, s1);
#else
                    pout[0] = s0; pout[1] = s1;
#endif
                } else if (v0) pout[0] = s0;
                if (edge_cta) {                           // uniform: only the first and last column blocks feed ghost columns
                    if (v0 && q0 < kHalo) pout[g.L] = s0;
                    if (v1 && q0 + 1 < kHalo) pout[1 + g.L] = s1;
                    if (v0 && q0 >= g.L - kHalo) pout[-g.L] = s0;
                    if (v1 && q0 + 1 >= g.L - kHalo) pout[1 - g.L] = s1;
                }
                Pu = Pc; Pc = Pd; Pd = Pn; Pl = Pln; Pr = Prn;
                pc = p1; p1 = p2; p2 += g.Lp;
                cell += (uint32_t)g.L;
            }
            pr2 += n_safe; grow += n_safe;
            if (g.wrap && pr2 == g.L) { pr2 = 0; p2 -= wrap_back; }
            if (grow == g.L) { grow = 0; cell = (uint32_t)qc; }
        }
        if (n_exact > 0) {   // second walk, extinction fringe only: the marked cells redone from their counters with the exact samplers
            for (int row = ra; row < rb; row++) {
                const int p = prow(row);
                int gr = g.row0 + row; if (gr < 0) gr += g.L; if (gr >= g.L) gr -= g.L;
                for (int k = 0; k < 2; k++) {
                    if (!(k ? v1 : v0)) continue;
                    double *o = A.pstar + (long long)trial * g.field_stride + (long long)p * g.Lp + col + k;
                    if (d2u(o[0]) != kExactMarker) continue;
                    const double *ic = A.in + (long long)trial * g.trial_stride + (long long)p * g.Lp + col + k;
                    const double up = ic[(long long)(prow(row - 1) - p) * g.Lp], dn = ic[(long long)(prow(row + 1) - p) * g.Lp];
                    const double alpha_i = A.dk.diffc * (STRICT ? up + dn + ic[-1] + ic[1] : (up + dn) + (ic[-1] + ic[1]));
                    const double v = riet_cell_exact(A.dk, (uint32_t)gr * (uint32_t)g.L + (uint32_t)(q0 + k), ptrial, A.step, A.seed, alpha_i, ic[0]);
                    o[0] = v;
                    if (q0 + k < kHalo) o[g.L] = v;
                    if (q0 + k >= g.L - kHalo) o[-g.L] = v;
                }
            }
        }
    }
}

} // namespace tml
