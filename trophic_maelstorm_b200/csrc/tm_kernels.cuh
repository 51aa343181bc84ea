// tm_kernels.cuh -- sm_100a kernels of the Dornic lattice step.
//
//   riet_step_kernel   Rietkerk SPB=1: vegetation Dornic (R/:3752-3809) + classical RK4 over
//                      (P,W,O) with f_2Dor (R/:2359-2440) + clamp/commit (R/:4150-4181) in ONE pass over HBM.
//                      (The <rho>_x / active-site reductions, R/:1211,1245, run as stats_kernel when a
//                      measurement is due: fusing them here cost 0.36 ms per measuring step in register
//                      spills against 0.08 ms for the separate 403 MB pass -- profiles/README.md.)
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
#ifndef TM_MAXTHREADS
#define TM_MAXTHREADS 160
#endif
#ifndef TM_MINBLOCKS
#define TM_MINBLOCKS 4
#endif
constexpr int kMaxThreads = TM_MAXTHREADS;  // CTA size cap (one thread per strip column)
constexpr int kMinBlocks = TM_MINBLOCKS;    // resident CTAs per SM the register budget is sized for

#ifndef TM_RHS_TWO_RCP
#define TM_RHS_TWO_RCP 1
#endif
#ifndef TM_STEP_ISSUERS
#define TM_STEP_ISSUERS 1
#endif
struct RietConsts {
    double dt, dt2, dt6;
    double cgmax, gmax, alpha, K1, K2, K2W0, rW;
    double Dp, Dw, Do;   // D[s]/dx^2
    double d;            // vegetation mortality (DET only; absorbed in beta otherwise)
    double rc;           // conversion c (DET: rc*gmax is recomputed as the reference does)
    double W0;
    double eps, lo;      // clamp
    double inv_lambda;   // 1/dk.lambda (FAST arithmetic only)
    double mu_scale;     // 2 (D0/dx^2) / sigma0^2 (FAST arithmetic only)
    double dt2cg, dtcg, dt6cg; // FAST, Dornic path: dt/2, dt, dt/6 times c*gmax (f_P = c gmax q is carried as q)
    double aK2W0, cP, cW, cO; // FAST: alpha*K2*W0, -(d+4Dp), -(rW+4Dw), -4Do
    double aKap;              // FAST: alpha*K2*(1-W0)
    DornicConsts dk;
};

struct Geom {
    int L, Lp, nrows, row0, wrap, rows_alloc, n_trials, n_fields;
    int ns, TX, nthreads;        // strips, output columns per strip, CTA size
    long long field_stride, trial_stride; // in doubles
};

// Philox4x32-10 key schedule: round r uses (k0 + r*0x9E3779B9, k1 + r*0xBB67AE85)
struct PhiloxKeys { uint32_t k[10][2]; };

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
    unsigned long long *counters; // [0] sampler/replay overflow, [1] pipeline timeouts
    double *dornic_out;           // DORNIC_ONLY: rho* [n_trials][nrows][L] (unpadded)
    const double *pstar;          // kSrcPre: rho* in the padded layout, [n_trials][rows_alloc][Lp]
    PhiloxKeys rk;                // the ten round keys of `seed` (host-computed)
    const double *logtab;         // device copy of the FAST-mode log table [128][2]
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
    if (mbar_try_wait(b, parity)) return true;
#pragma unroll 1
    for (int i = 0; i < (1 << 24); i++)
        if (mbar_try_wait(b, parity)) return true;
    return false;
}

// ------------------------------------------------------------------------------------------
// Lean FP64 helpers for the FAST (non-STRICT) arithmetic: hardware seed + Newton steps, no
// special-case branches, ~1 ulp.  Arguments are positive and far from the exponent limits.
// STRICT mode never uses them (IEEE division / sqrt, bit-identical to the oracle).
// ------------------------------------------------------------------------------------------
TM_D double rcp_lean(double b)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));   // MUFU.RCP64H, rel. error 2^-23
    r = fma(r, fma(-b, r, 1.0), r);
    r = fma(r, fma(-b, r, 1.0), r);
    return r;
#else
    return 1.0 / b;
#endif
}
// 1/b with ONE Newton step (rel. error ~2^-45): for the reaction terms only, whose result is scaled by dt
// before it meets the state (contribution to the state ~1e-17 relative).
TM_D double rcp_lean1(double b)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    return fma(r, fma(-b, r, 1.0), r);
#else
    return 1.0 / b;
#endif
}
TM_D double sqrt_lean(double a)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); // MUFU.RSQ64H, rel. error 2^-22
    const double h = 0.5 * a;
    y = fma(y, fma(-h * y, y, 0.5), y);                      // one Newton step on 1/sqrt(a): ~2^-43
    double s = a * y;
    s = fma(fma(-s, s, a), 0.5 * y, s);                      // Heron step on sqrt(a): error squared again, ~1 ulp
    return s;
#else
    return sqrt(a);
#endif
}
// sqrt(a) with ONE Newton step on the reciprocal root (rel. error ~1e-13): for the Gaussian Gamma shortcut
// only, where it multiplies a noise term that is at most 1/10 of the result (R/:3787, shape > 100).
TM_D double sqrt_lean1(double a)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    y = fma(y, fma(-h * y, y, 0.5), y);
    return a * y;
#else
    return sqrt(a);
#endif
}
// log(x) for x in (0,1] (the Box-Muller radius uniform), FAST arithmetic: x = 2^k m, m in [1,2); the top 7 mantissa
// bits pick c_i = 1 + (i + 1/2)/128 from a 128-entry table {1/c_i, log c_i}; r = m/c_i - 1 is below 2^-8 and
// log m = log c_i + (r - r^2/2 + r^3/3 - r^4/4 + r^5/5 - r^6/6).  No division; ~1e-16 absolute.
struct LogTab { double invc[128], logc[128]; };
TM_D double log_lean(double x, const double *tab /* [128][2] = {invc, logc} */)
{
    const double ln2 = TM_MC(0, 6.93147180559945286227e-01);
    const uint64_t b = d2u(x);
    const int k = (int)(b >> 52) - 1023;
    const int i = (int)((b >> 45) & 127);
    const double m = u2d((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    const double invc = tab[2 * i], logc = tab[2 * i + 1];
    const double r = fma(m, invc, -1.0);
    double p = fma(r, TM_MC(1, -1.0 / 6.0), TM_MC(2, 1.0 / 5.0));
    p = fma(r, p, -1.0 / 4.0);
    p = fma(r, p, TM_MC(3, 1.0 / 3.0));
    p = fma(r, p, -0.5);
    p = fma(r * r, p, r);
    return fma((double)k, ln2, logc) + p;
}

// ------------------------------------------------------------------------------------------
// Reaction terms
// ------------------------------------------------------------------------------------------
struct F3 { double p, w, o; };
struct alignas(16) WO { double w, o; };   // one column of a stage row: read and written as one 128-bit access

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
        // two MUFU-seeded reciprocals (one Newton step each): 10 FP64 operations for the two quotients where one shared
        // reciprocal of (W + K1)(P + K2) takes 12; alpha (P + K2 W0)/(P + K2) = alpha - alpha K2 (1 - W0)/(P + K2)
#if TM_RHS_TWO_RCP
        const double r1 = rcp_lean1(W + c.K1), r2 = rcp_lean1(P + c.K2);
        const double q = (W * P) * r1;                       // W P / (W + K1)
        const double infil = O * fma(-c.aKap, r2, c.alpha);
#else
        const double wk = W + c.K1, pk = P + c.K2;
        const double r = rcp_lean1(wk * pk);
        const double q = (W * P) * (pk * r);                 // W P / (W + K1)
        const double infil = fma(c.alpha, P, c.aK2W0) * (wk * r) * O;
#endif
        if (DET) f.p = fma(c.Dp, (Pu + Pd) + (Pl + Pr), fma(c.cP, P, c.cgmax * q));     // cP = -(d + 4 Dp)
        else     f.p = q;      // f_P / (c gmax): the factor is folded into the step constants dt2cg, dtcg, dt6cg
        f.w = fma(c.Dw, (Wu + Wd) + (Wl + Wr), fma(c.cW, W, fma(-c.gmax, q, infil))); // cW = -(rW + 4 Dw)
        f.o = fma(c.Do, (Ou + Od) + (Ol + Or_), fma(c.cO, O, a - infil));             // cO = -4 Do
    }
    return f;
}

template <bool STRICT>
TM_D double axpy(double a, double x, double y) { return STRICT ? (y + a * x) : fma(a, x, y); }

TM_D int hi32(double v) { return (int)(d2u(v) >> 32); }

// `if (v < eps && v >= lo) v = 0` (R/:4176): the first test almost never holds in live cells, so the
// second is only evaluated behind it.
TM_D double clamp_small(double v, double eps, double lo)
{
    if (v < eps) { if (v >= lo) v = 0.0; }
    return v;
}

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
        const uint64_t b = d2u(v), mag = b & 0x7FFFFFFFFFFFFFFFull;
        const bool nf = (b & 0x7FF0000000000000ull) == 0x7FF0000000000000ull;
        const bool isnan_ = mag > 0x7FF0000000000000ull;
        const bool neg = (b >> 63) != 0;
        active[f] += (!neg && mag != 0 && !isnan_) ? 1u : 0u;   // v > 0
        nonfinite += nf ? 1u : 0u;
        negative += (neg && mag != 0 && !isnan_) ? 1u : 0u;     // v < 0
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
// shared memory carve-up (doubles unless noted), RW = kMaxThreads + 4 (2 pad each side, 16 B aligned);
// the element of thread cx sits at [cx + 2] of a row:
//   pr   [kPSlots][RW]           TMA destination, P rows j-2 .. j+3 (stage-0 P = the Dornic input)
//   wo   [kWoSlots][2][RW]       TMA destination, W and O rows j-4 .. j+4
//   st   [3][kStSlots][NF][RW]   stage-1/2/3 states of the last kStSlots rows of each front: a row is [RW] x {W,O}
//                                interleaved (one 128-bit access per column), then [RW] x P when DET
//   mbar uint64[kWoSlots]        one per streamed row (its three bulk copies)
// The marching loop is unrolled by three rows, so a ring slot is (group base that rotates every three
// rows) + (compile-time offset): kPSlots, kWoSlots and kStSlots are multiples of three.
constexpr int kPSlots = 6;
constexpr int kWoSlots = 9;
constexpr int kStSlots = 3;
#ifndef TM_CARRY_U
#define TM_CARRY_U 1   // upper neighbours of the stage rows carried in registers instead of re-read from shared memory
#endif
#ifndef TM_CARRY_C
#define TM_CARRY_C 1   // ... and the centre values too: only the left/right neighbours come from shared memory
#endif
#ifndef TM_CARRY_0
#define TM_CARRY_0 0   // same for the stage-0 (streamed) W,O rows of front 1: 8 more registers spill at 96 (0.48 ms); at 128
#endif                 // registers / 3 CTAs per SM it is a wash (0.382 vs 0.379 ms per step), so it stays off
constexpr int kLead = 4;               // row j+kLead is requested when row j has been consumed

constexpr int kRW = kMaxThreads + 4;   // shared row width: compile-time, so field/neighbour offsets fold into immediates

TM_HD size_t riet_smem_bytes(int nthreads, bool det)
{
    const size_t RW = (size_t)kRW;
    const size_t nf = det ? 3 : 2;
    (void)nthreads;
    return sizeof(double) * RW * (kPSlots + 2 * kWoSlots + 3 * kStSlots * nf) + sizeof(uint64_t) * kWoSlots + 128;
}

// offset of row j+R (R in [-6,2]) in a ring made of groups of three slots: o0 = group of rows j..j+2,
// o1 = rows j-3..j-1, o2 = rows j-6..j-4 (group bases, in doubles); ss = doubles per slot
template <int R>
TM_D int ring_off(int o0, int o1, int o2, int ss)
{
    return (R >= 0) ? o0 + R * ss : (R >= -3) ? o1 + (R + 3) * ss : o2 + (R + 6) * ss;
}
template <int N> struct IntC { static constexpr int value = N; };

// The exact-sampler branches (Poisson mean <= 100, Gamma shape <= 100: extinction fringe only) are
// kept out of line so the marching loop stays small; they redo the cell from its counter, which
// yields the same value as the inline Gaussian path would have up to the point of divergence.
#if defined(__CUDACC__)
__device__ __noinline__
#else
inline
#endif
double riet_cell_exact(DornicConsts dk, uint32_t cell, uint32_t trial, uint32_t step, unsigned long long seed,
                       double alpha_i, double rho_i)
{
    PhiloxSrc src;
    src.begin(cell, trial, step, seed);
    return dornic_cell<kRietkerk>(dk, src, alpha_i, rho_i);
}

// The Box-Muller pair of a cell's Philox stream: the first two std::normal_distribution variates the
// generic normal01(PhiloxSrc) would return (uniforms 0 and 1 = Philox block 0), built straight from the
// block's bits.  Straight-line code, no rejection: every lane of a warp does the same work.
template <bool STRICT>
TM_D void gauss_pair(uint32_t cell, uint32_t trial, uint32_t step, const PhiloxKeys &rk, const double *logtab,
                     const double *sctab, double &n_first, double &n_second)
{
    uint32_t c0 = cell, c1 = trial, c2 = step, c3 = 0;
#pragma unroll
    for (int r = 0; r < 10; r++) { // same rounds as philox4x32_10(), keys read from the schedule
        uint32_t h0, l0, h1, l1;
        mulhilo32(0xD2511F53u, c0, h0, l0);
        mulhilo32(0xCD9E8D57u, c2, h1, l1);
        c0 = h1 ^ c1 ^ rk.k[r][0]; c1 = l1;
        c2 = h0 ^ c3 ^ rk.k[r][1]; c3 = l0;
    }
    // x = 2 - (1 + u0 | last bit) and a = 128 u1, each in one exact operation from the [1,2) / [128,256) mantissa doubles
    const double x = 2.0 - u2d(0x3FF0000000000001ull | ((((uint64_t)c1 << 32) | c0) >> 12));
    const double a = u2d(0x4060000000000000ull | ((((uint64_t)c3 << 32) | c2) >> 12)) - 128.0;
    double r;
    if (STRICT) r = sqrt(-2.0 * log_pos(x));
    else        r = sqrt_lean(fabs(-2.0 * log_lean(x, logtab)) + 0x1p-200); // (the bias keeps a cancelled log off rsqrt(0))
    double sn, cs;
    sincos_turn128(a, sctab, sn, cs);
    n_first = r * cs;
    n_second = r * sn;
}

// rho* of one cell of the Rietkerk vegetation field (R/:3756-3793) from a Philox stream, Gaussian regime only
// (Poisson mean > 100 and Gamma shape > 100: every cell away from the extinction fringe).  Cells that need the
// exact samplers come back with needs_exact = true and the caller runs riet_cell_exact() on them OUTSIDE its hot
// loop where it can: a call inside the loop makes the compiler keep the loop-carried state on the stack.
// nbsum = sum of the four neighbours; in_range = false for lanes outside the lattice.
template <bool STRICT>
TM_D double riet_dornic_value(const DornicConsts &dk, double inv_lambda, double mu_scale, double Pc, double nbsum, bool in_range,
                              uint32_t cell, uint32_t ptrial, uint32_t step, const PhiloxKeys &rk,
                              const double *logtab, const double *sctab, bool &needs_exact)
{
    const double alpha_i = dk.diffc * nbsum;
    const bool live = !((STRICT ? alpha_i : nbsum) == 0 && Pc == 0) && in_range;   // skip rule R/:3759
    const double ziggy = dk.ll * Pc;
    const double mu1 = STRICT ? (-1.0 + 2.0 * alpha_i * dk.sigma2_1) + 1.0  // mu + 1.0 in the reference's order
                              : mu_scale * nbsum;                            // 2 (D/dx^2) / sigma^2 folded on the host
    const bool gauss_n = live && ziggy > 100;
    const bool fastp = gauss_n || (live && ziggy == 0.0 && mu1 > 100);
    needs_exact = live && !fastp;
    double pstar = Pc;
    if (fastp) {
        double n_first, n_second;
        gauss_pair<STRICT>(cell, ptrial, step, rk, logtab, sctab, n_first, n_second);
        double gru = mu1, nz = n_first;
        if (STRICT) {
            if (gauss_n) { gru = mu1 + (double)(long long)fma(n_first, sqrt(ziggy), ziggy); nz = n_second; }   // R/:3776
            if (gru > 100) pstar = fma(nz, sqrt(gru), gru) / dk.lambda;
            else needs_exact = true;
        } else {
            // branch-free: both Gaussian shortcuts are evaluated and selected (a lane that takes neither discards a NaN/garbage
            // value); in the bulk every lane takes both, so the branches only cost convergence bookkeeping
            const double g2 = mu1 + trunc(fma(n_first, sqrt_lean(ziggy), ziggy));              // same value below 2^63
            gru = gauss_n ? g2 : mu1;
            nz = gauss_n ? n_second : n_first;
            const double v = fma(nz, sqrt_lean1(gru), gru) * inv_lambda;
            const bool ok = gru > 100;
            pstar = ok ? v : Pc;
            needs_exact = !ok;
        }
    }
    return pstar;
}

// SRC: where rho* comes from -- kSrcPhilox: sampled in front 1 (fused); kSrcReplay: same with host uniforms;
// kSrcPre: precomputed by riet_dornic_kernel into A.pstar (the split alternative, tm_set_fused(ctx, 0): the
// long serial sampler chain runs at full occupancy in its own kernel, this one keeps only the RK4 fronts).
enum { kSrcPhilox = 0, kSrcReplay = 1, kSrcPre = 2 };

template <bool DET, bool STRICT, int SRC, bool DORNIC_ONLY>
#if defined(__CUDACC__)
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks)
#else
void
#endif
riet_step_kernel(const StepArgs A)
{
    const Geom &g = A.g;
    const RietConsts &c = A.c;
    const int tid = threadIdx.x, nthreads = g.nthreads;
    constexpr int RW = kRW;
    constexpr int NF = DET ? 3 : 2;          // stage fields kept in shared memory: W,O[,P]
    constexpr int kSS = NF * RW;             // stage ring: doubles per slot (row)
    constexpr int kSG = kStSlots * kSS;      // ... per stage
    constexpr int kPGrp = 3 * RW;            // P ring: doubles per group of three rows
    constexpr int kWGrp = 3 * 2 * RW;        // W/O ring: doubles per group of three rows

    TM_DYN_SMEM(smem_raw);
    double *pr = reinterpret_cast<double *>(smem_raw);
    double *wo = pr + kPSlots * RW;
    double *st = wo + kWoSlots * 2 * RW;
    constexpr int kRingDoubles = (kPSlots + 2 * kWoSlots + 3 * kStSlots * NF) * RW;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(pr + kRingDoubles);
    const double *logtab = A.logtab; // 2 KB, read through L1 (shared memory is fully booked by the rings at 4 CTAs/SM)

    // TM_STEP_ISSUERS=3: a row's three bulk copies issued by three different threads (lane 0 of warps 0, 1, 2: rho*/P, W, O), each
    // with its own arrive.expect_tx on the row's mbarrier (count 3).  Measured slower than the single issuer (RK4 kernel 0.257 vs
    // 0.208 ms at L=4096, profiles/README.md round 2): the issue path lands on three warps' row bodies instead of one.  Default 1.
    const int n_issuers = (TM_STEP_ISSUERS == 3 && nthreads >= 96) ? 3 : 1;
    if (tid == 0) {
        for (int s = 0; s < kWoSlots; s++) mbar_init(&mbar[s], n_issuers);
        mbar_init_fence();
    }
    // rings start as zeros so that never-written slots are finite
    for (int i = tid; i < kRingDoubles; i += nthreads) pr[i] = 0.0;
    __syncthreads();

    // this CTA's share of the linearised (trial, strip, row) space
    const long long total = (long long)g.n_trials * g.ns * g.nrows;
    const long long w0 = total * blockIdx.x / gridDim.x;
    const long long w1 = total * (blockIdx.x + 1) / gridDim.x;
    const int TXH = g.TX + 2 * kHalo;                 // columns a strip touches
    const uint32_t row_bytes = (uint32_t)(((TXH + 1) & ~1) * sizeof(double));
    const int cx = tid;                               // column within the strip, halo included
    const bool col_in_strip = cx < TXH;
    const int eps_hi = hi32(c.eps);
    // step sizes of the P equation: in FAST Dornic mode rhs() returns f_P without its constant factor c*gmax
    const double hP2 = (STRICT || DET) ? c.dt2 : c.dt2cg, hP1 = (STRICT || DET) ? c.dt : c.dtcg, hP6 = (STRICT || DET) ? c.dt6 : c.dt6cg;

    // ring bookkeeping, advanced every three rows: group bases of rows j..j+2 / j-3..j-1 / j-6..j-4 in the
    // W/O ring (as group indices 0..2), of rows j..j+2 / j-3..j-1 in the P ring (0..1), and the mbarrier parity
    int G0 = 0, G1 = 2, G2 = 1, Q0 = 0, Q1 = 1;
    uint32_t par = 0;
    bool timeout = false;
    const double *const prc = pr + 2 + cx;
    const double *const woc = wo + 2 + cx;

    for (long long w = w0; w < w1;) {
        const int ts = (int)(w / g.nrows);            // trial*ns + strip
        const int trial = ts / g.ns, strip = ts % g.ns;
        const int ya = (int)(w % g.nrows);
        const int yb = (int)((w1 - w < (long long)(g.nrows - ya)) ? ya + (w1 - w) : g.nrows);
        w += yb - ya;

        const int x0 = strip * g.TX;                  // first output column of the strip
        const int gx = x0 - kHalo + cx;               // this thread's lattice column (may be <0 or >=L: ghost)
        int gxm = gx % g.L; if (gxm < 0) gxm += g.L;  // ... wrapped, for the Philox cell index
        const bool out_col = (cx >= kHalo) && (cx < kHalo + g.TX) && (gx < g.L);
        const bool ghost_r = out_col && gx < kHalo, ghost_l = out_col && gx >= g.L - kHalo; // also feeds a ghost column
        const double *inP = A.in + (long long)trial * g.trial_stride;
        const double *inW = inP + g.field_stride, *inO = inW + g.field_stride;
        const double *srcP = (SRC == kSrcPre) ? A.pstar + (long long)trial * g.field_stride : inP;
        double *outP = A.out + (long long)trial * g.trial_stride;
        const double a_rain = A.a_trial[trial];
        const int jlo = ya - kHalo, jhi = yb + kHalo - 1;
        const int nit = (jhi - jlo + 1 + 2) / 3 * 3;  // rows marched: padded to the unroll (<= 2 dummy rows, never stored)

        // physical row index of lattice-local row j in [-4, nrows+4)
        auto prow = [&](int j) -> long long {
            if (g.wrap) { int r = j % g.L; if (r < 0) r += g.L; return r; }
            return j + kHalo;
        };
        // Row streaming by thread 0: lattice-local row j (clamped to jhi: the dummy rows re-read it) goes into P-ring slot
        // `pslot` and W/O-ring slot `wslot` (= its mbarrier).  -4 <= j < nrows + 8, so the wrap needs no division.
        auto issue = [&](int j, int pslot, int wslot) {
            if ((tid & 31) != 0 || tid >= 32 * n_issuers) return;
            int r = (j < jhi) ? j : jhi;
            if (g.wrap) { if (r < 0) r += g.L; else if (r >= g.L) r -= g.L; } else r += kHalo;
            const long long off = (long long)r * g.Lp + x0;
            if (n_issuers == 1) {
                mbar_expect_tx(&mbar[wslot], 3 * row_bytes);
                bulk_g2s(pr + pslot * RW + 2, srcP + off, row_bytes, &mbar[wslot]);
                bulk_g2s(wo + wslot * 2 * RW + 2, inW + off, row_bytes, &mbar[wslot]);
                bulk_g2s(wo + wslot * 2 * RW + RW + 2, inO + off, row_bytes, &mbar[wslot]);
            } else {
                const int f = tid >> 5;   // 0: rho*/P, 1: W, 2: O
                mbar_expect_tx(&mbar[wslot], row_bytes);
                bulk_g2s(f == 0 ? pr + pslot * RW + 2 : wo + wslot * 2 * RW + (f - 1) * RW + 2,
                         (f == 0 ? srcP : f == 1 ? inW : inO) + off, row_bytes, &mbar[wslot]);
            }
        };
        issue(jlo, 3 * Q0 + 0, 3 * G0 + 0); issue(jlo + 1, 3 * Q0 + 1, 3 * G0 + 1); issue(jlo + 2, 3 * Q0 + 2, 3 * G0 + 2);
        if (nit > 3) issue(jlo + 3, 3 * Q1 + 0, 3 * G2 + 0);

        // global row (mod L) of front 1 as the Philox cell base, and the output pointer of the commit front
        int r1 = (g.row0 + jlo - 1) % g.L; if (r1 < 0) r1 += g.L;
        uint32_t cell_row = (uint32_t)r1 * (uint32_t)g.L;
        int r4 = (int)prow(jlo - 4);
        double *po_row = outP + ((long long)r4 * g.Lp + (gx + kHalo));

        // the few values that stay in registers (own column): rho* of rows j-4..j-2, the P stage
        // values, and the running K1+2K2+2K3 sums of the rows between fronts 1 and 4
        double Ps_m4 = 0, Ps_m3 = 0, Ps_m2 = 0;
        double P1_m2 = 0, P2_m3 = 0, P3_m4 = 0;
        double aP_m4 = 0, aP_m3 = 0, aP_m2 = 0, aW_m4 = 0, aW_m3 = 0, aW_m2 = 0, aO_m4 = 0, aO_m3 = 0, aO_m2 = 0;
        double w0_m1 = 0, o0_m1 = 0, w0_m2 = 0, o0_m2 = 0;   // stage-0 W,O of this column one and two rows behind the streamed row
        WO c1_m3 = {0, 0}, c2_m4 = {0, 0}, c3_m5 = {0, 0};   // own-column stage values one row behind each front (its "up")
        WO n1_m2 = {0, 0}, n2_m3 = {0, 0}, n3_m4 = {0, 0};   // ... and on each front's row (its centre): made by this thread one row ago

        for (int it = 0; it < nit; it += 3) {
            const int pq0 = Q0 * kPGrp, pq1 = Q1 * kPGrp;
            const int wg0 = G0 * kWGrp, wg1 = G1 * kWGrp, wg2 = G2 * kWGrp;

            auto body = [&](auto U) {
                constexpr int u = decltype(U)::value;
                const int j = jlo + it + u;
                if (!timeout && !mbar_wait(&mbar[3 * G0 + u], par)) timeout = true;
                // stage-0 rows of this column: P rows j-2..j, W/O rows j-4..j (W at [0], O at [RW])
                const double *p0 = prc + ring_off<u>(pq0, pq1, pq1, RW);
                const double *p1 = prc + ring_off<u - 1>(pq0, pq1, pq1, RW);
                const double *p2 = prc + ring_off<u - 2>(pq0, pq1, pq1, RW);
                const double *q0 = woc + ring_off<u>(wg0, wg1, wg2, 2 * RW);
                const double *q1 = woc + ring_off<u - 1>(wg0, wg1, wg2, 2 * RW);
                const double *q2 = woc + ring_off<u - 2>(wg0, wg1, wg2, 2 * RW);
                const double *q3 = woc + ring_off<u - 3>(wg0, wg1, wg2, 2 * RW);
                const double *q4 = woc + ring_off<u - 4>(wg0, wg1, wg2, 2 * RW);
                // stage rings: row j-back lives in slot (u - back) mod 3 (the segment starts on a multiple of 3)
                auto sp = [&](int stage, int back) -> WO * {       // {W,O} of this column
                    return reinterpret_cast<WO *>(st + (stage - 1) * kSG + (((u - back) % 3 + 3) % 3) * kSS) + 1 + cx;
                };
                auto spP = [&](int stage, int back) -> double * {  // P of this column (DET only)
                    return st + (stage - 1) * kSG + (((u - back) % 3 + 3) % 3) * kSS + 2 * RW + 2 + cx;
                };

                // ---- front 1: row j-1 : rho* and K1 -----------------------------------------------
                const double Pc = p1[0], Pu = p2[0], Pd = p0[0], Pl = p1[-1], Pr = p1[1];
                double pstar = Pc;
                if (!DET && SRC != kSrcPre) {
                    const uint32_t cell = cell_row + (uint32_t)gxm;
                    if (SRC == kSrcReplay) {
                        const double alpha_i = c.dk.diffc * (Pu + Pd + Pl + Pr);
                        if (!(alpha_i == 0 && Pc == 0) && col_in_strip) { // skip rule R/:3759
                            ReplaySrc src;
                            src.begin(A.replay_u + ((size_t)trial * g.L * g.L + cell) * A.stride, A.stride);
                            pstar = dornic_cell<kRietkerk>(c.dk, src, alpha_i, Pc);
                            if (src.overflow_count() && out_col && j - 1 >= ya && j - 1 < yb)
                                atomicAdd(&A.counters[0], (unsigned long long)src.overflow_count());
                        }
                    } else {
                        const double nbsum = STRICT ? Pu + Pd + Pl + Pr : (Pu + Pd) + (Pl + Pr); // STRICT: the reference's order R/:3756
                        bool needs_exact;
                        pstar = riet_dornic_value<STRICT>(c.dk, c.inv_lambda, c.mu_scale, Pc, nbsum, col_in_strip,
                                                          cell, A.trial0 + (uint32_t)trial, A.step, A.rk, logtab, sincos_table(), needs_exact);
                        if (needs_exact) // extinction fringe only (the pipeline needs the value now, so the call stays in the loop here)
                            pstar = riet_cell_exact(c.dk, cell, A.trial0 + (uint32_t)trial, A.step, A.seed, c.dk.diffc * nbsum, Pc);
                    }
                }
                if (DORNIC_ONLY) {
                    if (out_col && j - 1 >= ya && j - 1 < yb)
                        A.dornic_out[((size_t)trial * g.nrows + (j - 1)) * g.L + gx] = pstar;
                } else {
                    // stage-0 W,O of this column: the row below is read, centre and upper row were read one and two rows ago
                    const double W0d = q0[0], O0d = q0[RW];
                    const double W0c = TM_CARRY_0 ? w0_m1 : q1[0], O0c = TM_CARRY_0 ? o0_m1 : q1[RW];
                    const double W0u = TM_CARRY_0 ? w0_m2 : q2[0], O0u = TM_CARRY_0 ? o0_m2 : q2[RW];
                    const F3 k1 = rhs<DET, STRICT>(c, a_rain, pstar, W0c, O0c,
                                                   W0u, W0d, q1[-1], q1[1],
                                                   O0u, O0d, q1[RW - 1], q1[RW + 1],
                                                   Pu, Pd, Pl, Pr);
                    const double P1_m1 = axpy<STRICT>(hP2, k1.p, pstar);
                    const double W1_m1 = axpy<STRICT>(c.dt2, k1.w, W0c);
                    const double O1_m1 = axpy<STRICT>(c.dt2, k1.o, O0c);
                    *sp(1, 1) = WO{W1_m1, O1_m1};
                    if (DET) *spP(1, 1) = P1_m1;

                    // ---- front 2: row j-2 : K2 (stage-1 rows j-3, j-2 from the ring, j-1 just made) ---
                    const WO *s1 = sp(1, 2);
                    const WO s1c = TM_CARRY_C ? n1_m2 : s1[0], s1l = s1[-1], s1r = s1[1], s1u = TM_CARRY_U ? c1_m3 : *sp(1, 3);
                    const F3 k2 = rhs<DET, STRICT>(c, a_rain, P1_m2, s1c.w, s1c.o,
                                                   s1u.w, W1_m1, s1l.w, s1r.w,
                                                   s1u.o, O1_m1, s1l.o, s1r.o,
                                                   DET ? *spP(1, 3) : 0.0, P1_m1, DET ? spP(1, 2)[-1] : 0.0, DET ? spP(1, 2)[1] : 0.0);
                    const double P2_m2 = axpy<STRICT>(hP2, k2.p, Ps_m2);
                    const double W2_m2 = axpy<STRICT>(c.dt2, k2.w, W0u);
                    const double O2_m2 = axpy<STRICT>(c.dt2, k2.o, O0u);
                    *sp(2, 2) = WO{W2_m2, O2_m2};
                    if (DET) *spP(2, 2) = P2_m2;
                    aP_m2 = axpy<STRICT>(2.0, k2.p, aP_m2); aW_m2 = axpy<STRICT>(2.0, k2.w, aW_m2); aO_m2 = axpy<STRICT>(2.0, k2.o, aO_m2);

                    // ---- front 3: row j-3 : K3 ---------------------------------------------------------
                    const WO *s2 = sp(2, 3);
                    const WO s2c = TM_CARRY_C ? n2_m3 : s2[0], s2l = s2[-1], s2r = s2[1], s2u = TM_CARRY_U ? c2_m4 : *sp(2, 4);
                    const F3 k3 = rhs<DET, STRICT>(c, a_rain, P2_m3, s2c.w, s2c.o,
                                                   s2u.w, W2_m2, s2l.w, s2r.w,
                                                   s2u.o, O2_m2, s2l.o, s2r.o,
                                                   DET ? *spP(2, 4) : 0.0, P2_m2, DET ? spP(2, 3)[-1] : 0.0, DET ? spP(2, 3)[1] : 0.0);
                    const double P3_m3 = axpy<STRICT>(hP1, k3.p, Ps_m3);
                    const double W3_m3 = axpy<STRICT>(c.dt, k3.w, q3[0]);
                    const double O3_m3 = axpy<STRICT>(c.dt, k3.o, q3[RW]);
                    *sp(3, 3) = WO{W3_m3, O3_m3};
                    if (DET) *spP(3, 3) = P3_m3;
                    aP_m3 = axpy<STRICT>(2.0, k3.p, aP_m3); aW_m3 = axpy<STRICT>(2.0, k3.w, aW_m3); aO_m3 = axpy<STRICT>(2.0, k3.o, aO_m3);

                    // ---- front 4: row j-4 : K4, commit -------------------------------------------------
                    const WO *s3 = sp(3, 4);
                    const WO s3c = TM_CARRY_C ? n3_m4 : s3[0], s3l = s3[-1], s3r = s3[1], s3u = TM_CARRY_U ? c3_m5 : *sp(3, 5);
                    const F3 k4 = rhs<DET, STRICT>(c, a_rain, P3_m4, s3c.w, s3c.o,
                                                   s3u.w, W3_m3, s3l.w, s3r.w,
                                                   s3u.o, O3_m3, s3l.o, s3r.o,
                                                   DET ? *spP(3, 5) : 0.0, P3_m3, DET ? spP(3, 4)[-1] : 0.0, DET ? spP(3, 4)[1] : 0.0);
                    double nP = axpy<STRICT>(hP6, aP_m4 + k4.p, Ps_m4);
                    double nW = axpy<STRICT>(c.dt6, aW_m4 + k4.w, q4[0]);
                    double nO = axpy<STRICT>(c.dt6, aO_m4 + k4.o, q4[RW]);
                    // R/:4176 `if (v < eps && v >= lo) v = 0`, almost never taken for live cells: pre-tested on the high words
                    // (integer compares; v < eps implies hi(v) <= hi(eps), and a set sign bit compares below)
                    if (!DET && (hi32(nP) <= eps_hi || hi32(nW) <= eps_hi || hi32(nO) <= eps_hi)) {
                        nP = clamp_small(nP, c.eps, c.lo); nW = clamp_small(nW, c.eps, c.lo); nO = clamp_small(nO, c.eps, c.lo);
                    }
                    if (out_col && j - 4 >= ya && j - 4 < yb) {
                        double *po = po_row; // row r4, padded-layout column gx + 4
                        po[0] = nP; po[g.field_stride] = nW; po[2 * g.field_stride] = nO;
                        if (ghost_r) { po[g.L] = nP; po[g.field_stride + g.L] = nW; po[2 * g.field_stride + g.L] = nO; }
                        if (ghost_l) { po[-g.L] = nP; po[g.field_stride - g.L] = nW; po[2 * g.field_stride - g.L] = nO; }
                    }
                    // ---- shift the register windows by one row (renamed away by the unrolling) -------------
                    w0_m2 = W0c; o0_m2 = O0c; w0_m1 = W0d; o0_m1 = O0d;
                    c1_m3 = s1c; c2_m4 = s2c; c3_m5 = s3c;   // this row's centre is the next row's upper neighbour
                    n1_m2 = WO{W1_m1, O1_m1}; n2_m3 = WO{W2_m2, O2_m2}; n3_m4 = WO{W3_m3, O3_m3}; // ... and what was just made, its centre
                    P3_m4 = P3_m3; P2_m3 = P2_m2; P1_m2 = P1_m1;
                    aP_m4 = aP_m3; aP_m3 = aP_m2; aP_m2 = k1.p; aW_m4 = aW_m3; aW_m3 = aW_m2; aW_m2 = k1.w;
                    aO_m4 = aO_m3; aO_m3 = aO_m2; aO_m2 = k1.o;
                    Ps_m4 = Ps_m3; Ps_m3 = Ps_m2; Ps_m2 = pstar;
                }
                if (++r1 == g.L) { r1 = 0; cell_row = 0; } else cell_row += (uint32_t)g.L;
                if (g.wrap && r4 + 1 == g.L) { r4 = 0; po_row -= (long long)(g.L - 1) * g.Lp; } else { r4 += 1; po_row += g.Lp; }

                __syncthreads();
                // row j has been consumed by everyone: request row j+kLead into the slots of rows j-5 (W,O) and j-2 (P)
                if (it + u + kLead < nit) {
                    if (u == 0) issue(j + kLead, 3 * Q1 + 1, 3 * G2 + 1);
                    else if (u == 1) issue(j + kLead, 3 * Q1 + 2, 3 * G2 + 2);
                    else issue(j + kLead, 3 * Q0 + 0, 3 * G1 + 0);
                }
            };
            body(IntC<0>()); body(IntC<1>()); body(IntC<2>());

            { const int t = G0; G0 = G2; G2 = G1; G1 = t; }   // next three rows go where rows j-6..j-4 were
            { const int t = Q0; Q0 = Q1; Q1 = t; }
            if (G0 == 0) par ^= 1u;
        }
    }
    if (timeout && tid == 0) atomicAdd(&A.counters[1], 1ull);
}

// ------------------------------------------------------------------------------------------
// Rietkerk vegetation Dornic sweep as its own kernel (the default pairing: this sweep, then the RK4 fronts
// of riet_step_kernel<.., kSrcPre, ..>; tm_set_fused(ctx, 1) selects the single fused kernel instead)
// ------------------------------------------------------------------------------------------
// One thread per column, marching down a run of rows: the centre and upper neighbour roll through registers,
// the row below and the two horizontal neighbours are plain coalesced loads (L1 hits after the first touch of
// a line), so a cell costs one HBM read and one HBM write.  No pipeline state beyond that: 64 registers,
// 32 warps per SM, and the long serial chain (Philox -> log -> sqrt -> sincos -> trunc -> sqrt) of one warp
// hides behind seven others.  Writes rho* in the padded layout (ghost columns included).  The pair moves 64 B per
// site-update instead of the fused kernel's 48 B, on kernels that are bound by FP64/issue slots, not by HBM.
struct DornicArgs {
    Geom g;
    DornicConsts dk;
    double inv_lambda, mu_scale;   // FAST arithmetic: 1/lambda, 2 (D0/dx^2)/sigma0^2
    const double *in;              // state buffer (field 0 = P is read)
    double *pstar;                 // [n_trials][rows_alloc][Lp]
    unsigned long long seed;
    unsigned int trial0, step;
    PhiloxKeys rk;
    int row_lo, row_hi;            // lattice-local rows [row_lo, row_hi) to sample (slabs: -3 .. nrows+3)
    int rows_per_cta;              // rows one CTA marches through
    const double *logtab;          // device copy of the FAST-mode log table [128][2]
};

#ifndef TM_DORNIC_MINBLOCKS
#define TM_DORNIC_MINBLOCKS 4
#endif
#ifndef TM_DORNIC_UNROLL
#define TM_DORNIC_UNROLL 4
#endif
constexpr int kDornicUnroll = TM_DORNIC_UNROLL;
constexpr int kDornicThreads = 256;
constexpr uint64_t kExactMarker = 0x7FF8C0DEC0DE0001ull; // a NaN no computation produces: 'redo this cell exactly'
TM_HD size_t riet_dornic_smem_bytes() { return sizeof(double) * 512; } // log table + sin/cos table

template <bool STRICT>
#if defined(__CUDACC__)
__global__ void __launch_bounds__(kDornicThreads, TM_DORNIC_MINBLOCKS)
#else
void
#endif
riet_dornic_kernel(const DornicArgs A)
{
    const Geom &g = A.g;
    const int tid = threadIdx.x;
    TM_DYN_SMEM(smem_raw);
    double *logtab = reinterpret_cast<double *>(smem_raw); // [128][2]
    double *sctab = logtab + 256;                          // [128][2]
    if (!STRICT) logtab[tid] = A.logtab[tid];
    sctab[tid] = sincos_table()[tid];
    __syncthreads();
    // work item = (trial, column block of 256, run of rows_per_cta rows)
    const int ncb = (g.L + kDornicThreads - 1) / kDornicThreads;
    const int nrow = A.row_hi - A.row_lo;
    const int nrun = (nrow + A.rows_per_cta - 1) / A.rows_per_cta;
    const long long items = (long long)g.n_trials * ncb * nrun;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int run = (int)(it % nrun);
        const int cb = (int)((it / nrun) % ncb);
        const int trial = (int)(it / ((long long)nrun * ncb));
        const int ra = A.row_lo + run * A.rows_per_cta;
        const int rb = (ra + A.rows_per_cta < A.row_hi) ? ra + A.rows_per_cta : A.row_hi;
        const int q = cb * kDornicThreads + tid;
        const bool valid = q < g.L;
        const int qc = valid ? q : g.L - 1;          // clamped: idle lanes read a legal address
        const bool ghost_r = valid && q < kHalo, ghost_l = valid && q >= g.L - kHalo;
        const bool edge_cta = (cb == 0) || (cb == ncb - 1);
        const uint32_t ptrial = A.trial0 + (uint32_t)trial;
        // physical row of lattice-local row r in [-4, nrows+4): whole lattices wrap, slabs carry ghost rows
        auto prow = [&](int r) -> int {
            if (g.wrap) { if (r < 0) r += g.L; if (r >= g.L) r -= g.L; return r; }
            return r + kHalo;
        };
        const long long col = (long long)kHalo + qc;
        const double *const inT = A.in + (long long)trial * g.trial_stride + col;       // (physical row 0, column qc) of P
        // rho* has the same row/column layout as a field: its element sits at a fixed offset from the input element
        const long long odelta = (A.pstar + (long long)trial * g.field_stride) - (A.in + (long long)trial * g.trial_stride);
        // Three row pointers walk down the column: pc -> row, p1 -> row+1, p2 -> row+2 (physical rows, periodic wrap
        // applied to p2 only, once per row).  The centre column rolls through registers two rows ahead: the load of
        // row+2 is in flight for a whole iteration (the L2/HBM latency of the row stream is what this kernel would
        // otherwise wait for).  The loop is unrolled by two so the pointer and window rotations are renamed away.
        const double *pc = inT + (long long)prow(ra) * g.Lp, *p1 = inT + (long long)prow(ra + 1) * g.Lp;
        int pr2 = prow(ra + 2);
        const double *p2 = inT + (long long)pr2 * g.Lp;
        const long long wrap_back = (long long)g.L * g.Lp;
        double Pu = inT[(long long)prow(ra - 1) * g.Lp], Pc = pc[0], Pd = p1[0];
        double Pl = pc[-1], Pr = pc[1];                                      // ghost columns cover q = 0 and q = L-1
        int grow = g.row0 + ra; if (grow < 0) grow += g.L; if (grow >= g.L) grow -= g.L;
        uint32_t cell = (uint32_t)grow * (uint32_t)g.L + (uint32_t)qc;
        int n_exact = 0;
        // The periodic wraps (of the prefetch pointer: row+2 reaching physical row L; of the global row that keys the Philox
        // cell) are uniform events that happen at most once or twice per run: the rows between them are marched by an inner
        // loop with no wrap tests at all, the events are applied between inner loops.
        for (int row = ra; row < rb;) {
            int n_safe = rb - row;
            if (g.wrap && g.L - pr2 < n_safe) n_safe = g.L - pr2;
            if (g.L - grow < n_safe) n_safe = g.L - grow;
            const int row_end = row + n_safe;
#pragma unroll (kDornicUnroll)
            for (; row < row_end; row++) {
                double Pn = Pd, Pln = Pl, Prn = Pr;
                if (row + 1 < rb) { Pn = p2[0]; Pln = p1[-1]; Prn = p1[1]; }     // for the next iteration
                const double nbsum = STRICT ? Pu + Pd + Pl + Pr : (Pu + Pd) + (Pl + Pr);    // STRICT: the reference's order R/:3756
                bool needs_exact;
                double pstar = riet_dornic_value<STRICT>(A.dk, A.inv_lambda, A.mu_scale, Pc, nbsum, valid,
                                                         cell, ptrial, A.step, A.rk, logtab, sctab, needs_exact);
                if (needs_exact) { pstar = u2d(kExactMarker); n_exact++; }    // resolved by the second walk below
                if (valid) {
                    double *pout = const_cast<double *>(pc) + odelta;
                    pout[0] = pstar;
                    if (edge_cta) {                                   // uniform: only the first and last column blocks feed ghost columns
                        if (ghost_r) pout[g.L] = pstar;
                        if (ghost_l) pout[-g.L] = pstar;
                    }
                }
                Pu = Pc; Pc = Pd; Pd = Pn; Pl = Pln; Pr = Prn;
                pc = p1; p1 = p2; p2 += g.Lp;
                cell += (uint32_t)g.L;
            }
            pr2 += n_safe; grow += n_safe;
            if (g.wrap && pr2 == g.L) { pr2 = 0; p2 -= wrap_back; }
            if (grow == g.L) { grow = 0; cell = (uint32_t)qc; }
        }
        // second walk, extinction fringe only: the cells of this thread's run that need the exact Poisson / Gamma
        // samplers (marked above) are redone from their counters, out of the hot loop
        if (n_exact > 0) {
            for (int row = ra; row < rb; row++) {
                const int p = prow(row);
                double *o = A.pstar + (long long)trial * g.field_stride + (long long)p * g.Lp + col;
                if (d2u(o[0]) != kExactMarker) continue;
                const double *ic = A.in + (long long)trial * g.trial_stride + (long long)p * g.Lp + col;
                const double up = ic[(long long)(prow(row - 1) - p) * g.Lp], dn = ic[(long long)(prow(row + 1) - p) * g.Lp];
                int gr = g.row0 + row; if (gr < 0) gr += g.L; if (gr >= g.L) gr -= g.L;
                const double alpha_i = A.dk.diffc * (STRICT ? up + dn + ic[-1] + ic[1] : (up + dn) + (ic[-1] + ic[1])); // as in the first walk
                const double v = riet_cell_exact(A.dk, (uint32_t)gr * (uint32_t)g.L + (uint32_t)qc, ptrial, A.step, A.seed, alpha_i, ic[0]);
                o[0] = v;
                if (ghost_r) o[g.L] = v;
                if (ghost_l) o[-g.L] = v;
            }
        }
    }
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
    // whole rows per block (coalesced along q, no per-cell division)
    for (int r = blk; r < g.nrows; r += A.blocks_per_trial)
    for (int q = tid; q < g.L; q += nthreads) {
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
    for (int f = 0; f < g.n_fields; f++) {
        const double *in = A.in + (long long)trial * g.trial_stride + (long long)f * g.field_stride + kHalo;
        for (int r = blk; r < g.nrows; r += A.blocks_per_trial) { // whole rows per block: coalesced, no per-element division
            const double *row = in + (long long)(g.wrap ? r : r + kHalo) * g.Lp;
            for (int q = tid; q < g.L; q += nthreads) tstat.add(f, row[q]);
        }
    }
    flush_stats(tstat, g.n_fields, trial, A.partials + blockIdx.x, scratch, tid, nthreads);
}

struct TmStatsDev { double mean[3]; unsigned long long active[3], nonfinite, negative; };

// one block per trial: thread t adds the records t, t+nthreads, ... in index order, then the per-thread
// partials are added in thread order -- deterministic for a given launch geometry
#if defined(__CUDACC__)
__global__ void
#else
inline void
#endif
stats_finalize_kernel(const PartialStats *partials, int n_records, int n_trials, double cells, TmStatsDev *out)
{
    const int t = blockIdx.x, tid = threadIdx.x, nthreads = blockDim.x;
    if (t >= n_trials) return;
    TM_DYN_SMEM(smem_raw);
    double *sh = reinterpret_cast<double *>(smem_raw);                 // [3][nthreads]
    unsigned long long *shu = reinterpret_cast<unsigned long long *>(sh + 3 * nthreads); // [5][nthreads]
    double sum[3] = {0.0, 0.0, 0.0};
    unsigned long long cnt[5] = {0, 0, 0, 0, 0};
    for (int i = tid; i < n_records; i += nthreads) {
        if (partials[i].trial != t) continue;
        for (int f = 0; f < 3; f++) { sum[f] += partials[i].sum[f]; cnt[f] += partials[i].active[f]; }
        cnt[3] += partials[i].nonfinite; cnt[4] += partials[i].negative;
    }
    for (int f = 0; f < 3; f++) sh[f * nthreads + tid] = sum[f];
    for (int k = 0; k < 5; k++) shu[k * nthreads + tid] = cnt[k];
    __syncthreads();
    if (tid == 0) {
        TmStatsDev s;
        for (int f = 0; f < 3; f++) {
            double a = 0.0; unsigned long long n = 0;
            for (int i = 0; i < nthreads; i++) { a += sh[f * nthreads + i]; n += shu[f * nthreads + i]; }
            s.mean[f] = a / cells; // sum/(double)size, R/:1218
            s.active[f] = n;
        }
        unsigned long long nf = 0, ng = 0;
        for (int i = 0; i < nthreads; i++) { nf += shu[3 * nthreads + i]; ng += shu[4 * nthreads + i]; }
        s.nonfinite = nf; s.negative = ng;
        out[t] = s;
    }
}

// Spreading-experiment observables of field 0 (P/:156-169 meansq_spread_of_vector, P/:380-390 theborderwall):
// per trial the number of active sites N, sum over them of (p-c_x)^2 + (q-c_y)^2 (p = row, q = column), and
// whether any site of the four border lines is active.  All three are integers (the squared distances are exact
// in the reference's double sum too, far below 2^53), so 64-bit atomics give an order-independent, exact result.
struct SpreadArgs { Geom g; const double *in; unsigned long long *acc; /* [n_trials][3] = N, sum r^2, border */ int c_x, c_y, blocks_per_trial; };
#if defined(__CUDACC__)
__global__ void __launch_bounds__(256)
#else
inline void
#endif
spread_kernel(const SpreadArgs A)
{
    const Geom &g = A.g;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int trial = blockIdx.x / A.blocks_per_trial, blk = blockIdx.x % A.blocks_per_trial;
    const double *in = A.in + (long long)trial * g.trial_stride + kHalo;
    unsigned long long n = 0, sq = 0, border = 0;
    for (int r = blk; r < g.nrows; r += A.blocks_per_trial) {
        const double *row = in + (long long)(g.wrap ? r : r + kHalo) * g.Lp;
        const long long p = g.row0 + r, dp = p - A.c_x;
        for (int q = tid; q < g.L; q += nthreads) {
            if (row[q] > 0) {
                const long long dq = q - A.c_y;
                n += 1; sq += (unsigned long long)(dp * dp + dq * dq);
                if (p == 0 || p == g.L - 1 || q == 0 || q == g.L - 1) border = 1;
            }
        }
    }
    if (n) { atomicAdd(&A.acc[3 * trial], n); atomicAdd(&A.acc[3 * trial + 1], sq); }
    if (border) atomicAdd(&A.acc[3 * trial + 2], 1ull);
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
