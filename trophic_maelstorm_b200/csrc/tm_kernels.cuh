// tm_kernels.cuh -- sm_100a kernels of the Dornic lattice step.
//
//   riet_step_kernel   Rietkerk SPB=1: vegetation Dornic (R/:3752-3809) + classical RK4 over
//                      (P,W,O) with f_2Dor (R/:2359-2440) + clamp/commit (R/:4150-4181) + the
//                      <rho>_x / active-site reductions (R/:1211,1245), all in ONE pass over HBM.
//                      DET=true is the noise-off path: RK4 with the full RHS f_2D (R/:1718-1814).
//   dp_step_kernel     DP / cubic DP: Dornic (P/:1032-1061) + closed form (P/:1064-1068) or
//                      pointwise RK4 (P/:1775-1790) + reductions (P/:123,145).
//   stats_kernel, stats_finalize_kernel   stand-alone reductions of a resident state.
//
// HBM layout (one allocation per ping-pong side):
//   double [n_trials][n_fields][rows_alloc][Lp]
//   a row holds [4 ghost cols = cells L-4..L-1][L cells][4 ghost cols = cells 0..3][pad to Lp];
//   Lp is a multiple of 16 doubles.  Whole lattices have rows_alloc = L and wrap rows by index
//   arithmetic; row slabs (domain decomposition) have rows_alloc = nrows + 8 with 4 ghost rows
//   above and below, filled by the halo exchange.
//
// The Rietkerk kernel is a persistent, marching, register-pipelined stencil.  A CTA owns a
// column strip of TX output columns (+4 halo columns each side, one thread per column) and a
// contiguous run of rows.  Input rows are streamed into a shared-memory ring by TMA bulk copies
// (cp.async.bulk + mbarrier).  At pipeline step j a thread holds, for its column, the rolling
// vertical windows of every RK stage in registers and advances four fronts at once:
//     row j-1: rho* (Dornic) and K1      row j-2: K2      row j-3: K3      row j-4: K4, commit
// Horizontal neighbours of the stage states are exchanged through double-buffered shared rows
// (one __syncthreads per step).  Every input element is read from HBM once and every output
// written once: 48 B per site-update for the three fields.  Halo columns/rows are recomputed;
// because each cell's random numbers come from a Philox counter keyed by its GLOBAL index,
// recomputed cells draw identical values, so the result does not depend on the decomposition.
#pragma once
#include "tm_samplers.cuh"

#if defined(__CUDACC__)
#define TM_D __device__ __forceinline__
#define TM_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#else
#include "tm_cpu_emu.h" // tests/emu only: runs the same source on host threads
#define TM_D inline
#endif

namespace tml {

constexpr int kHalo = 4;            // RK4 dependency cone: 3 stencil applications + the Dornic alpha stencil
constexpr int kRing = 8;            // TMA input ring slots
constexpr int kPrefetch = kRing - 2;
constexpr int kMaxThreads = 384;

struct RietConsts {
    double dt, dt2, dt6;
    double cgmax, gmax, alpha, K1, K2, K2W0, rW;
    double Dp, Dw, Do;   // D[s]/dx^2
    double d;            // vegetation mortality (DET only; absorbed in beta otherwise)
    double rc;           // conversion c (DET: rc*gmax is recomputed as the reference does)
    double W0;
    double eps, lo;      // clamp
    DornicConsts dk;
};

struct Geom {
    int L, Lp, nrows, row0, wrap, rows_alloc, n_trials, n_fields;
    int ns, TX, nthreads;        // strips, output columns per strip, CTA size
    long long field_stride, trial_stride; // in doubles
};

struct PartialStats { // one record per (CTA, segment)
    int trial, pad;
    double sum[3];
    unsigned long long active[3], nonfinite, negative;
};

struct StepArgs {
    Geom g;
    RietConsts c;
    const double *in;
    double *out;
    const double *a_trial;        // [n_trials] control parameter
    unsigned long long seed;
    unsigned int trial0, step;
    const double *replay_u;       // REPLAY: [n_trials][L*L][stride] (global cell order)
    unsigned int stride;
    PartialStats *partials;       // [gridDim.x][max_segs]
    int max_segs;
    unsigned long long *counters; // [0] sampler/replay overflow, [1] pipeline timeouts
    double *dornic_out;           // DORNIC_ONLY: rho* [n_trials][nrows][L] (unpadded)
};

// ------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (sm_90+/sm_100a PTX); host emulation in tm_cpu_emu.h
// ------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
TM_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
TM_D void mbar_init(uint64_t *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
TM_D void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TM_D void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
TM_D void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
TM_D bool mbar_try_wait(uint64_t *b, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
#endif

// bounded wait: a lost copy must not hang the GPU; it is counted and the step's output is garbage
TM_D bool mbar_wait(uint64_t *b, uint32_t parity)
{
    for (int i = 0; i < (1 << 24); i++)
        if (mbar_try_wait(b, parity)) return true;
    return false;
}

// ------------------------------------------------------------------------------------------
// Reaction terms
// ------------------------------------------------------------------------------------------
struct F3 { double p, w, o; };

// f_2Dor R/:2376-2382 (DET=false) / f_2D R/:1730-1737 (DET=true); STRICT keeps the reference's
// operation order so results are bit-identical to the oracle; the fast form shares one reciprocal.
template <bool DET, bool STRICT>
TM_D F3 rhs(const RietConsts &c, double a, double P, double W, double O,
            double Wu, double Wd, double Wl, double Wr, double Ou, double Od, double Ol, double Or_,
            double Pu, double Pd, double Pl, double Pr)
{
    F3 f;
    if (STRICT) {
        if (DET) {
            f.p = c.rc * c.gmax * W * P / (W + c.K1) - c.d * P + c.Dp * (Pu + Pd + Pl + Pr - 4 * P);
        } else {
            f.p = c.cgmax * W * P / (W + c.K1);
        }
        f.w = c.alpha * (P + c.K2 * c.W0) / (P + c.K2) * O - c.gmax * W * P / (W + c.K1) - c.rW * W
            + c.Dw * (Wu + Wd + Wl + Wr - 4 * W);
        f.o = a - c.alpha * (P + c.K2 * c.W0) / (P + c.K2) * O + c.Do * (Ou + Od + Ol + Or_ - 4 * O);
    } else {
        const double wk = W + c.K1, pk = P + c.K2;
        const double r = 1.0 / (wk * pk);
        const double q = (W * P) * (pk * r);                 // W P / (W + K1)
        const double infil = c.alpha * (P + c.K2W0) * (wk * r) * O;
        const double lapW = fma(-4.0, W, (Wu + Wd) + (Wl + Wr));
        const double lapO = fma(-4.0, O, (Ou + Od) + (Ol + Or_));
        f.p = c.cgmax * q;
        if (DET) {
            const double lapP = fma(-4.0, P, (Pu + Pd) + (Pl + Pr));
            f.p = fma(c.Dp, lapP, fma(-c.d, P, f.p));
        }
        f.w = fma(c.Dw, lapW, fma(-c.rW, W, fma(-c.gmax, q, infil)));
        f.o = fma(c.Do, lapO, a - infil);
    }
    return f;
}

template <bool STRICT>
TM_D double axpy(double a, double x, double y) { return STRICT ? (y + a * x) : fma(a, x, y); }

TM_D double clamp_small(double v, double eps, double lo) { return (v < eps && v >= lo) ? 0.0 : v; }

// ------------------------------------------------------------------------------------------
// Block-level statistics flush: sum/active/guards of this thread's outputs -> one record
// ------------------------------------------------------------------------------------------
struct ThreadStats {
    double sum[3];
    unsigned int active[3], nonfinite, negative;
    TM_D void clear() { for (int f = 0; f < 3; f++) { sum[f] = 0.0; active[f] = 0; } nonfinite = 0; negative = 0; }
    TM_D void add(int f, double v)
    {
        sum[f] += v;
        const uint64_t b = d2u(v);
        active[f] += (v > 0) ? 1u : 0u;
        nonfinite += ((b & 0x7FF0000000000000ull) == 0x7FF0000000000000ull) ? 1u : 0u;
        negative += (v < 0) ? 1u : 0u;
    }
};

// scratch: double[8][nthreads]
TM_D void flush_stats(const ThreadStats &ts, int n_fields, int trial, PartialStats *rec, double *scratch, int tid, int nthreads)
{
    for (int f = 0; f < 3; f++) {
        scratch[f * nthreads + tid] = ts.sum[f];
        scratch[(3 + f) * nthreads + tid] = (double)ts.active[f];
    }
    scratch[6 * nthreads + tid] = (double)ts.nonfinite;
    scratch[7 * nthreads + tid] = (double)ts.negative;
    __syncthreads();
    if (tid < 8) {
        double s = 0.0; // fixed order: deterministic for a given launch geometry
        for (int i = 0; i < nthreads; i++) s += scratch[tid * nthreads + i];
        if (tid < 3) rec->sum[tid] = (tid < n_fields) ? s : 0.0;
        else if (tid < 6) rec->active[tid - 3] = (unsigned long long)s;
        else if (tid == 6) rec->nonfinite = (unsigned long long)s;
        else rec->negative = (unsigned long long)s;
        if (tid == 0) { rec->trial = trial; rec->pad = 0; }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Rietkerk fused step
// ------------------------------------------------------------------------------------------
// shared memory carve-up (doubles unless noted), RW = nthreads + 4 (2 pad each side, 16 B aligned):
//   ring  [kRing][3][RW]      TMA destination rows (P,W,O); element of thread cx at [cx + 2]
//   exch  [2][NX][RW]         stage states published for horizontal neighbours, NX = 6 (W1,O1,W2,O2,W3,O3) or 9 (+P1,P2,P3)
//   scratch [8][nthreads]     statistics flush
//   mbar  uint64[kRing]
TM_HD size_t riet_smem_bytes(int nthreads, bool det)
{
    const size_t RW = (size_t)nthreads + 4;
    return sizeof(double) * (kRing * 3 * RW + 2 * (det ? 9 : 6) * RW + 8 * (size_t)nthreads) + sizeof(uint64_t) * kRing + 128;
}

template <bool DET, bool STRICT, bool REPLAY, bool STATS, bool DORNIC_ONLY>
#if defined(__CUDACC__)
__global__ void __launch_bounds__(kMaxThreads, 1)
#else
void
#endif
riet_step_kernel(const StepArgs A)
{
    const Geom &g = A.g;
    const RietConsts &c = A.c;
    const int tid = threadIdx.x, nthreads = g.nthreads;
    const int RW = nthreads + 4;
    constexpr int NX = DET ? 9 : 6;

    TM_DYN_SMEM(smem_raw);
    double *ring = reinterpret_cast<double *>(smem_raw);
    double *exch = ring + kRing * 3 * RW;
    double *scratch = exch + 2 * NX * RW;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(scratch + 8 * nthreads);

    if (tid == 0) {
        for (int s = 0; s < kRing; s++) mbar_init(&mbar[s], 1);
        mbar_init_fence();
    }
    // exchange rows start as zeros so that never-published halo slots are at least finite
    for (int i = tid; i < 2 * NX * RW; i += nthreads) exch[i] = 0.0;
    __syncthreads();

    // this CTA's share of the linearised (trial, strip, row) space
    const long long total = (long long)g.n_trials * g.ns * g.nrows;
    const long long w0 = total * blockIdx.x / gridDim.x;
    const long long w1 = total * (blockIdx.x + 1) / gridDim.x;
    const int TXH = g.TX + 2 * kHalo;                 // columns a strip touches
    const uint32_t row_bytes = (uint32_t)(((TXH + 1) & ~1) * sizeof(double));
    const bool active_thread = tid < TXH;
    const int cx = tid;                               // column within the strip, halo included

    long long seq = 0;   // rows consumed so far (ring slot / phase bookkeeping)
    int seg_no = 0;
    bool timeout = false;

    for (long long w = w0; w < w1;) {
        const int ts = (int)(w / g.nrows);            // trial*ns + strip
        const int trial = ts / g.ns, strip = ts % g.ns;
        const int ya = (int)(w % g.nrows);
        const int yb = (int)((w1 - w < (long long)(g.nrows - ya)) ? ya + (w1 - w) : g.nrows);
        w += yb - ya;

        const int x0 = strip * g.TX;                  // first output column of the strip
        const int gx = x0 - kHalo + cx;               // this thread's lattice column (may be <0 or >=L: ghost)
        const int xs = x0;                            // padded-layout column of cx = 0 is (x0 - 4) + 4
        const double *inP = A.in + (long long)trial * g.trial_stride;
        const double *inW = inP + g.field_stride, *inO = inW + g.field_stride;
        double *outP = A.out + (long long)trial * g.trial_stride;
        double *outW = outP + g.field_stride, *outO = outW + g.field_stride;
        const double a_rain = A.a_trial[trial];
        const int jlo = ya - kHalo, jhi = yb + kHalo - 1;

        // physical row index of lattice-local row j in [-4, nrows+4)
        auto prow = [&](int j) -> long long {
            if (g.wrap) { int r = j % g.L; if (r < 0) r += g.L; return r; }
            return j + kHalo;
        };
        auto issue = [&](int j, long long n) { // thread 0: stream row j into ring slot n % kRing
            const int slot = (int)(n % kRing);
            double *dst = ring + (size_t)slot * 3 * RW + 2;
            const long long off = prow(j) * g.Lp + xs;
            mbar_expect_tx(&mbar[slot], 3 * row_bytes);
            bulk_g2s(dst, inP + off, row_bytes, &mbar[slot]);
            bulk_g2s(dst + RW, inW + off, row_bytes, &mbar[slot]);
            bulk_g2s(dst + 2 * RW, inO + off, row_bytes, &mbar[slot]);
        };
        if (tid == 0)
            for (int k = 0; k < kPrefetch && jlo + k <= jhi; k++) issue(jlo + k, seq + k);

        // rolling register windows (own column); suffix = row relative to j
        double Pp_m2 = 0, Pp_m1 = 0;
        double W0_m4 = 0, W0_m3 = 0, W0_m2 = 0, W0_m1 = 0, O0_m4 = 0, O0_m3 = 0, O0_m2 = 0, O0_m1 = 0;
        double Ps_m4 = 0, Ps_m3 = 0, Ps_m2 = 0;           // rho* (or P in DET)
        double W1_m3 = 0, W1_m2 = 0, O1_m3 = 0, O1_m2 = 0, P1_m2 = 0, P1_m3 = 0;
        double W2_m4 = 0, W2_m3 = 0, O2_m4 = 0, O2_m3 = 0, P2_m3 = 0, P2_m4 = 0;
        double W3_m5 = 0, W3_m4 = 0, O3_m5 = 0, O3_m4 = 0, P3_m4 = 0, P3_m5 = 0;
        double aP_m4 = 0, aP_m3 = 0, aP_m2 = 0, aW_m4 = 0, aW_m3 = 0, aW_m2 = 0, aO_m4 = 0, aO_m3 = 0, aO_m2 = 0;
        ThreadStats tstat; tstat.clear();

        for (int j = jlo; j <= jhi; j++, seq++) {
            const int slot = (int)(seq % kRing), pslot = (int)((seq + kRing - 1) % kRing);
            if (!timeout && !mbar_wait(&mbar[slot], (uint32_t)((seq / kRing) & 1))) timeout = true;
            const double *rj = ring + (size_t)slot * 3 * RW + 2 + cx;   // row j, this column
            const double *rp = ring + (size_t)pslot * 3 * RW + 2 + cx;  // row j-1
            const double *xr = exch + (size_t)((seq + 1) & 1) * NX * RW + 2 + cx; // published at step j-1
            double *xw = exch + (size_t)(seq & 1) * NX * RW + 2 + cx;

            const bool en1 = (j - 1 >= ya - 3) && (j - 1 <= yb + 2);
            const bool en2 = (j - 2 >= ya - 2) && (j - 2 <= yb + 1);
            const bool en3 = (j - 3 >= ya - 1) && (j - 3 <= yb);
            const bool en4 = (j - 4 >= ya) && (j - 4 <= yb - 1);

            double Pn = 0, Wn = 0, On = 0;
            if (active_thread) { Pn = rj[0]; Wn = rj[RW]; On = rj[2 * RW]; }

            // ---- front 1: row j-1 : rho* and K1 -----------------------------------------------
            double Ps_m1 = 0, W1_m1 = 0, O1_m1 = 0, P1_m1 = 0, aP_m1 = 0, aW_m1 = 0, aO_m1 = 0;
            if (en1 && active_thread) {
                const double pl = rp[-1], pr = rp[1];
                double pstar = Pp_m1;
                if (!DET) {
                    const double alpha_i = c.dk.diffc * (Pp_m2 + Pn + pl + pr);
                    if (!(alpha_i == 0 && Pp_m1 == 0)) {
                        int lr = g.row0 + (j - 1); // global row of this cell
                        lr %= g.L; if (lr < 0) lr += g.L;
                        int lc = gx % g.L; if (lc < 0) lc += g.L;
                        const uint32_t cell = (uint32_t)lr * (uint32_t)g.L + (uint32_t)lc;
                        if (REPLAY) {
                            ReplaySrc src;
                            src.begin(A.replay_u + ((size_t)trial * g.L * g.L + cell) * A.stride, A.stride);
                            pstar = dornic_cell<kRietkerk>(c.dk, src, alpha_i, Pp_m1);
                            if (src.overflow_count() && cx >= kHalo && cx < kHalo + g.TX && gx < g.L && j - 1 >= ya && j - 1 < yb)
                                atomicAdd(&A.counters[0], (unsigned long long)src.overflow_count());
                        } else {
                            PhiloxSrc src;
                            src.begin(cell, A.trial0 + (uint32_t)trial, A.step, A.seed);
                            pstar = dornic_cell<kRietkerk>(c.dk, src, alpha_i, Pp_m1);
                        }
                    }
                }
                Ps_m1 = pstar;
                if (!DORNIC_ONLY) {
                    const F3 k1 = rhs<DET, STRICT>(c, a_rain, pstar, W0_m1, O0_m1,
                                                   W0_m2, Wn, rp[RW - 1], rp[RW + 1],
                                                   O0_m2, On, rp[2 * RW - 1], rp[2 * RW + 1],
                                                   Pp_m2, Pn, pl, pr);
                    P1_m1 = axpy<STRICT>(c.dt2, k1.p, pstar);
                    W1_m1 = axpy<STRICT>(c.dt2, k1.w, W0_m1);
                    O1_m1 = axpy<STRICT>(c.dt2, k1.o, O0_m1);
                    aP_m1 = k1.p; aW_m1 = k1.w; aO_m1 = k1.o;
                    xw[0] = W1_m1; xw[RW] = O1_m1;
                    if (DET) xw[6 * RW] = P1_m1;
                }
            }
            if (DORNIC_ONLY) {
                if (en1 && active_thread && cx >= kHalo && cx < kHalo + g.TX && gx < g.L && j - 1 >= ya && j - 1 < yb)
                    A.dornic_out[((size_t)trial * g.nrows + (j - 1)) * g.L + gx] = Ps_m1;
            } else {
                // ---- front 2: row j-2 : K2 ---------------------------------------------------
                double W2_m2 = 0, O2_m2 = 0, P2_m2 = 0;
                if (en2 && active_thread) {
                    const F3 k2 = rhs<DET, STRICT>(c, a_rain, P1_m2, W1_m2, O1_m2,
                                                   W1_m3, W1_m1, xr[-1], xr[1],
                                                   O1_m3, O1_m1, xr[RW - 1], xr[RW + 1],
                                                   P1_m3, P1_m1, DET ? xr[6 * RW - 1] : 0.0, DET ? xr[6 * RW + 1] : 0.0);
                    P2_m2 = axpy<STRICT>(c.dt2, k2.p, Ps_m2);
                    W2_m2 = axpy<STRICT>(c.dt2, k2.w, W0_m2);
                    O2_m2 = axpy<STRICT>(c.dt2, k2.o, O0_m2);
                    aP_m2 = axpy<STRICT>(2.0, k2.p, aP_m2); aW_m2 = axpy<STRICT>(2.0, k2.w, aW_m2); aO_m2 = axpy<STRICT>(2.0, k2.o, aO_m2);
                    xw[2 * RW] = W2_m2; xw[3 * RW] = O2_m2;
                    if (DET) xw[7 * RW] = P2_m2;
                }
                // ---- front 3: row j-3 : K3 ---------------------------------------------------
                double W3_m3 = 0, O3_m3 = 0, P3_m3 = 0;
                if (en3 && active_thread) {
                    const F3 k3 = rhs<DET, STRICT>(c, a_rain, P2_m3, W2_m3, O2_m3,
                                                   W2_m4, W2_m2, xr[2 * RW - 1], xr[2 * RW + 1],
                                                   O2_m4, O2_m2, xr[3 * RW - 1], xr[3 * RW + 1],
                                                   P2_m4, P2_m2, DET ? xr[7 * RW - 1] : 0.0, DET ? xr[7 * RW + 1] : 0.0);
                    P3_m3 = axpy<STRICT>(c.dt, k3.p, Ps_m3);
                    W3_m3 = axpy<STRICT>(c.dt, k3.w, W0_m3);
                    O3_m3 = axpy<STRICT>(c.dt, k3.o, O0_m3);
                    aP_m3 = axpy<STRICT>(2.0, k3.p, aP_m3); aW_m3 = axpy<STRICT>(2.0, k3.w, aW_m3); aO_m3 = axpy<STRICT>(2.0, k3.o, aO_m3);
                    xw[4 * RW] = W3_m3; xw[5 * RW] = O3_m3;
                    if (DET) xw[8 * RW] = P3_m3;
                }
                // ---- front 4: row j-4 : K4, commit -------------------------------------------
                if (en4 && active_thread) {
                    const F3 k4 = rhs<DET, STRICT>(c, a_rain, P3_m4, W3_m4, O3_m4,
                                                   W3_m5, W3_m3, xr[4 * RW - 1], xr[4 * RW + 1],
                                                   O3_m5, O3_m3, xr[5 * RW - 1], xr[5 * RW + 1],
                                                   P3_m5, P3_m3, DET ? xr[8 * RW - 1] : 0.0, DET ? xr[8 * RW + 1] : 0.0);
                    double nP = axpy<STRICT>(c.dt6, aP_m4 + k4.p, Ps_m4);
                    double nW = axpy<STRICT>(c.dt6, aW_m4 + k4.w, W0_m4);
                    double nO = axpy<STRICT>(c.dt6, aO_m4 + k4.o, O0_m4);
                    if (!DET) { nP = clamp_small(nP, c.eps, c.lo); nW = clamp_small(nW, c.eps, c.lo); nO = clamp_small(nO, c.eps, c.lo); }
                    if (cx >= kHalo && cx < kHalo + g.TX && gx < g.L) {
                        const long long ro = prow(j - 4) * g.Lp;
                        const int pc = gx + kHalo; // padded-layout column
                        outP[ro + pc] = nP; outW[ro + pc] = nW; outO[ro + pc] = nO;
                        if (gx < kHalo) { outP[ro + pc + g.L] = nP; outW[ro + pc + g.L] = nW; outO[ro + pc + g.L] = nO; }
                        if (gx >= g.L - kHalo) { outP[ro + pc - g.L] = nP; outW[ro + pc - g.L] = nW; outO[ro + pc - g.L] = nO; }
                        if (STATS) { tstat.add(0, nP); tstat.add(1, nW); tstat.add(2, nO); }
                    }
                }
                // ---- shift the windows by one row ---------------------------------------------
                W3_m5 = W3_m4; W3_m4 = W3_m3; O3_m5 = O3_m4; O3_m4 = O3_m3; P3_m5 = P3_m4; P3_m4 = P3_m3;
                W2_m4 = W2_m3; W2_m3 = W2_m2; O2_m4 = O2_m3; O2_m3 = O2_m2; P2_m4 = P2_m3; P2_m3 = P2_m2;
                W1_m3 = W1_m2; W1_m2 = W1_m1; O1_m3 = O1_m2; O1_m2 = O1_m1; P1_m3 = P1_m2; P1_m2 = P1_m1;
                aP_m4 = aP_m3; aP_m3 = aP_m2; aP_m2 = aP_m1; aW_m4 = aW_m3; aW_m3 = aW_m2; aW_m2 = aW_m1;
                aO_m4 = aO_m3; aO_m3 = aO_m2; aO_m2 = aO_m1;
                Ps_m4 = Ps_m3; Ps_m3 = Ps_m2; Ps_m2 = Ps_m1;
                W0_m4 = W0_m3; W0_m3 = W0_m2; O0_m4 = O0_m3; O0_m3 = O0_m2;
            }
            W0_m2 = W0_m1; W0_m1 = Wn; O0_m2 = O0_m1; O0_m1 = On;
            Pp_m2 = Pp_m1; Pp_m1 = Pn;

            __syncthreads();
            if (tid == 0 && j + kPrefetch <= jhi) issue(j + kPrefetch, seq + kPrefetch);
        }

        if (STATS && !DORNIC_ONLY) {
            PartialStats *rec = A.partials + (size_t)blockIdx.x * A.max_segs + (seg_no < A.max_segs ? seg_no : A.max_segs - 1);
            flush_stats(tstat, 3, trial, rec, scratch, tid, nthreads);
        }
        seg_no++;
    }
    if (timeout && tid == 0) atomicAdd(&A.counters[1], 1ull);
}

// ------------------------------------------------------------------------------------------
// DP / cubic DP step: one thread per cell, 5-point reads served by L1/L2
// ------------------------------------------------------------------------------------------
struct DpConsts { double dt, dt2, dt6, b, c; };
struct DpTrial { DornicConsts dk; }; // per trial: beta, lambda depend on the control parameter a

struct DpArgs {
    Geom g;
    DpConsts c;
    const double *in;
    double *out;
    const DpTrial *trials;        // [n_trials]
    unsigned long long seed;
    unsigned int trial0, step;
    const double *replay_u;
    unsigned int stride;
    PartialStats *partials;       // [gridDim.x]  (a block never straddles trials)
    unsigned long long *counters;
    double *dornic_out;
    int blocks_per_trial;
};

template <int MODEL, bool REPLAY, bool STATS, bool DORNIC_ONLY>
#if defined(__CUDACC__)
__global__ void __launch_bounds__(256)
#else
void
#endif
dp_step_kernel(const DpArgs A)
{
    const Geom &g = A.g;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int trial = blockIdx.x / A.blocks_per_trial;
    const int blk = blockIdx.x % A.blocks_per_trial;
    TM_DYN_SMEM(smem_raw);
    double *scratch = reinterpret_cast<double *>(smem_raw);
    const double *in = A.in + (long long)trial * g.trial_stride;
    double *out = A.out + (long long)trial * g.trial_stride;
    const DornicConsts dk = A.trials[trial].dk;
    ThreadStats tstat; tstat.clear();
    const long long ncell = (long long)g.nrows * g.L;
    for (long long i = (long long)blk * nthreads + tid; i < ncell; i += (long long)A.blocks_per_trial * nthreads) {
        const int r = (int)(i / g.L), q = (int)(i % g.L);
        long long ru, rd;
        if (g.wrap) { ru = (r == 0) ? g.L - 1 : r - 1; rd = (r == g.L - 1) ? 0 : r + 1; }
        else        { ru = r - 1 + kHalo; rd = r + 1 + kHalo; }
        const long long rc = g.wrap ? r : r + kHalo;
        const int pc = q + kHalo;
        const double up = in[ru * g.Lp + pc], dn = in[rd * g.Lp + pc];
        const double lf = in[rc * g.Lp + pc - 1], rt = in[rc * g.Lp + pc + 1];
        const double ctr = in[rc * g.Lp + pc];
        const double alpha_i = dk.diffc * (up + dn + lf + rt);
        double star = ctr;
        if (!(alpha_i == 0 && ctr == 0)) {
            int gr = (g.row0 + r) % g.L;
            const uint32_t cell = (uint32_t)gr * (uint32_t)g.L + (uint32_t)q;
            if (REPLAY) {
                ReplaySrc src;
                src.begin(A.replay_u + ((size_t)trial * g.L * g.L + cell) * A.stride, A.stride);
                star = dornic_cell<MODEL>(dk, src, alpha_i, ctr);
                if (src.overflow_count()) atomicAdd(&A.counters[0], (unsigned long long)src.overflow_count());
            } else {
                PhiloxSrc src;
                src.begin(cell, A.trial0 + (uint32_t)trial, A.step, A.seed);
                star = dornic_cell<MODEL>(dk, src, alpha_i, ctr);
            }
        }
        if (DORNIC_ONLY) {
            A.dornic_out[((size_t)trial * g.nrows + r) * g.L + q] = star;
            continue;
        }
        double v;
        if (MODEL == kDP) {
            v = star / (1 + A.c.b * star * A.c.dt); // P/:1066
        } else { // RK4_Integrate P/:1784-1790, reference operation order
            const double b = A.c.b, c = A.c.c, dt = A.c.dt, dt2 = A.c.dt2, dt6 = A.c.dt6;
            double K1, K2, K3, K4, M;
            K1 = -b * star * star - c * star * star * star;
            M = star + dt2 * K1;
            K2 = -b * M * M - c * M * M * M; M = star + dt2 * K2;
            K3 = -b * M * M - c * M * M * M; M = star + dt * K3;
            K4 = -b * M * M - c * M * M * M;
            v = star + dt6 * (K1 + 2.0 * K2 + 2.0 * K3 + K4);
        }
        out[rc * g.Lp + pc] = v;
        if (q < kHalo) out[rc * g.Lp + pc + g.L] = v;
        if (q >= g.L - kHalo) out[rc * g.Lp + pc - g.L] = v;
        if (STATS) tstat.add(0, v);
    }
    if (STATS && !DORNIC_ONLY)
        flush_stats(tstat, 1, trial, A.partials + blockIdx.x, scratch, tid, nthreads);
}

// ------------------------------------------------------------------------------------------
// Stand-alone statistics of a resident state, and the deterministic finalisation
// ------------------------------------------------------------------------------------------
struct StatsArgs {
    Geom g;
    const double *in;
    PartialStats *partials; // [gridDim.x]
    int blocks_per_trial;
};

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256)
#else
inline void
#endif
stats_kernel(const StatsArgs A)
{
    const Geom &g = A.g;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int trial = blockIdx.x / A.blocks_per_trial, blk = blockIdx.x % A.blocks_per_trial;
    TM_DYN_SMEM(smem_raw);
    double *scratch = reinterpret_cast<double *>(smem_raw);
    ThreadStats tstat; tstat.clear();
    const long long ncell = (long long)g.nrows * g.L;
    for (int f = 0; f < g.n_fields; f++) {
        const double *in = A.in + (long long)trial * g.trial_stride + (long long)f * g.field_stride;
        for (long long i = (long long)blk * nthreads + tid; i < ncell; i += (long long)A.blocks_per_trial * nthreads) {
            const int r = (int)(i / g.L), q = (int)(i % g.L);
            const long long rc = g.wrap ? r : r + kHalo;
            tstat.add(f, in[rc * g.Lp + q + kHalo]);
        }
    }
    flush_stats(tstat, g.n_fields, trial, A.partials + blockIdx.x, scratch, tid, nthreads);
}

struct TmStatsDev { double mean[3]; unsigned long long active[3], nonfinite, negative; };

// one thread per trial sums the partial records in index order
#if defined(__CUDACC__)
__global__ void
#else
inline void
#endif
stats_finalize_kernel(const PartialStats *partials, int n_records, int n_trials, double cells, TmStatsDev *out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trials) return;
    TmStatsDev s;
    for (int f = 0; f < 3; f++) { s.mean[f] = 0.0; s.active[f] = 0; }
    s.nonfinite = 0; s.negative = 0;
    for (int i = 0; i < n_records; i++) {
        if (partials[i].trial != t) continue;
        for (int f = 0; f < 3; f++) { s.mean[f] += partials[i].sum[f]; s.active[f] += partials[i].active[f]; }
        s.nonfinite += partials[i].nonfinite; s.negative += partials[i].negative;
    }
    for (int f = 0; f < 3; f++) s.mean[f] /= cells; // sum/(double)size, R/:1218
    out[t] = s;
}

// ghost-column fill after an upload / halo unpack: one thread per (row, 0..7)
struct GhostArgs { Geom g; double *buf; };
#if defined(__CUDACC__)
__global__ void
#else
inline void
#endif
ghost_cols_kernel(const GhostArgs A)
{
    const Geom &g = A.g;
    const long long n = (long long)g.n_trials * g.n_fields * g.rows_alloc * 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % 8);
        const long long row = i / 8; // flat over trials, fields, rows
        double *r = A.buf + row * g.Lp;
        if (k < 4) r[k] = r[g.L + k];                 // left ghosts  = cells L-4..L-1
        else       r[g.L + kHalo + (k - 4)] = r[kHalo + (k - 4)]; // right ghosts = cells 0..3
    }
}

} // namespace tml
