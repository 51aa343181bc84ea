// tm_math.cuh -- elementary functions, Philox and uniforms of the sampler spec.
//
// Everything here is built from IEEE-754 +,-,*,/,sqrt, explicit fma and integer bit operations,
// and the translation unit is compiled with --fmad=false, so results are a pure function of the
// inputs: identical on every sm_100a launch geometry and on a host compile (tests/emu).
//
// Reference call sites these serve: the std:: samplers of P/:1055-1059 and R/:3776-3793
// (libstdc++ <bits/random.tcc>), the engine std::mt19937_64 of P/:975, R/:3513.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define TM_HD __host__ __device__ __forceinline__
#else
#define TM_HD inline
#endif

namespace tml {

TM_HD uint64_t d2u(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; __builtin_memcpy(&u, &x, 8); return u;
#endif
}
TM_HD double u2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; __builtin_memcpy(&x, &u, 8); return x;
#endif
}

// Rejection loops are bounded (spec): never reached with genuine uniforms, but a starved replay
// buffer or a garbage halo cell must not spin.
constexpr int kMaxReject = 64;
constexpr int kMaxProduct = 4096;

// log(x), finite normal x > 0: x = 2^k m, m in [sqrt(1/2), sqrt(2)), log m = 2 atanh(f/(2+f)).
// Full-width FP64 constants of log_pos / exp_any.  As literals each use costs two move instructions on the device (a 64-bit
// immediate only fits an FP64 instruction when its low word is zero); read from the constant bank they are plain operands.
// TM_KC(table, index, literal): the table entry on the device, the literal on the host -- the same bits (each table is
// initialised from the very same literals).
#define TM_LOG_CONSTS 6.93147180369123816490e-01, 1.90821492927058770002e-10, 6.666666666666735130e-01, 3.999999999940941908e-01, \
    2.857142874366239149e-01, 2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01
#define TM_EXP_CONSTS 1.44269504088896338700e+00, 1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, \
    1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0
#if defined(__CUDACC__)
static __constant__ double kLogC[9] = { TM_LOG_CONSTS };
static __constant__ double kExpC[12] = { TM_EXP_CONSTS };
#endif
#if defined(__CUDA_ARCH__)
#define TM_KC(tab, i, lit) tab[i]
#else
#define TM_KC(tab, i, lit) (lit)
#endif

TM_HD double log_pos(double x)
{
    const double ln2_hi = TM_KC(kLogC, 0, 6.93147180369123816490e-01), ln2_lo = TM_KC(kLogC, 1, 1.90821492927058770002e-10);
    const double Lg1 = TM_KC(kLogC, 2, 6.666666666666735130e-01), Lg2 = TM_KC(kLogC, 3, 3.999999999940941908e-01),
                 Lg3 = TM_KC(kLogC, 4, 2.857142874366239149e-01), Lg4 = TM_KC(kLogC, 5, 2.222219843214978396e-01),
                 Lg5 = TM_KC(kLogC, 6, 1.818357216161805012e-01), Lg6 = TM_KC(kLogC, 7, 1.531383769920937332e-01),
                 Lg7 = TM_KC(kLogC, 8, 1.479819860511658591e-01);
    if (!(x > 0.0)) return (x == 0.0) ? -INFINITY : NAN;
    if (x == INFINITY) return x;
    uint64_t b = d2u(x);
    int k = (int)(b >> 52) - 1023;
    const uint64_t mant = b & 0x000FFFFFFFFFFFFFull;
    if (mant > 0x6A09E667F3BCCull) { b = mant | 0x3FE0000000000000ull; k += 1; }
    else                           { b = mant | 0x3FF0000000000000ull; }
    const double f = u2d(b) - 1.0;
    const double dk = (double)k;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
    const double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    return dk * ln2_hi - ((hfsq - fma(s, hfsq + R, dk * ln2_lo)) - f);
}

// exp(x): k = rint(x log2 e), two-part reduction, degree-13 Taylor, two-step 2^k scaling.
TM_HD double exp_any(double x)
{
    const double LOG2E = TM_KC(kExpC, 0, 1.44269504088896338700e+00);
    const double ln2_hi = TM_KC(kLogC, 0, 6.93147180369123816490e-01), ln2_lo = TM_KC(kLogC, 1, 1.90821492927058770002e-10);
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.2) return 0.0;
    const double kf = rint(x * LOG2E);
    double r = fma(kf, -ln2_hi, x);
    r = fma(kf, -ln2_lo, r);
    double q = TM_KC(kExpC, 1, 1.0 / 6227020800.0);
    q = fma(q, r, TM_KC(kExpC, 2, 1.0 / 479001600.0));
    q = fma(q, r, TM_KC(kExpC, 3, 1.0 / 39916800.0));
    q = fma(q, r, TM_KC(kExpC, 4, 1.0 / 3628800.0));
    q = fma(q, r, TM_KC(kExpC, 5, 1.0 / 362880.0));
    q = fma(q, r, TM_KC(kExpC, 6, 1.0 / 40320.0));
    q = fma(q, r, TM_KC(kExpC, 7, 1.0 / 5040.0));
    q = fma(q, r, TM_KC(kExpC, 8, 1.0 / 720.0));
    q = fma(q, r, TM_KC(kExpC, 9, 1.0 / 120.0));
    q = fma(q, r, TM_KC(kExpC, 10, 1.0 / 24.0));
    q = fma(q, r, TM_KC(kExpC, 11, 1.0 / 6.0));
    q = fma(q, r, 0.5);
    double y = 1.0 + fma(r * r, q, r);
    const int k = (int)kf;
    const int k1 = k / 2, k2 = k - k1;
    y *= u2d((uint64_t)(k1 + 1023) << 52);
    y *= u2d((uint64_t)(k2 + 1023) << 52);
    return y;
}

// ln((k-1)!) for k = 1..40, correctly rounded (60-digit decimal arithmetic; these are the same 40
// mathematical constants the test oracle generates for itself with oracle/gen_tables.py).
TM_HD double logfact_small(int k)
{
    switch (k) {
    case 1: return 0.0; case 2: return 0.0;
    case 3: return 0.6931471805599453;  case 4: return 1.791759469228055;
    case 5: return 3.1780538303479458;  case 6: return 4.787491742782046;
    case 7: return 6.579251212010101;   case 8: return 8.525161361065415;
    case 9: return 10.60460290274525;   case 10: return 12.801827480081469;
    case 11: return 15.104412573075516; case 12: return 17.502307845873887;
    case 13: return 19.987214495661885; case 14: return 22.552163853123425;
    case 15: return 25.19122118273868;  case 16: return 27.89927138384089;
    case 17: return 30.671860106080672; case 18: return 33.50507345013689;
    case 19: return 36.39544520803305;  case 20: return 39.339884187199495;
    case 21: return 42.335616460753485; case 22: return 45.38013889847691;
    case 23: return 48.47118135183523;  case 24: return 51.60667556776438;
    case 25: return 54.78472939811232;  case 26: return 58.00360522298052;
    case 27: return 61.261701761002;    case 28: return 64.55753862700634;
    case 29: return 67.88974313718154;  case 30: return 71.25703896716801;
    case 31: return 74.65823634883016;  case 32: return 78.0922235533153;
    case 33: return 81.55795945611504;  case 34: return 85.05446701758152;
    case 35: return 88.58082754219768;  case 36: return 92.1361756036871;
    case 37: return 95.7196945421432;   case 38: return 99.33061245478743;
    case 39: return 102.96819861451381; default: return 106.63176026064346;
    }
}

// log Gamma(k) for integer-valued k >= 1 (LS:1280, LS:1388 only ever pass such arguments).
TM_HD double lgamma_int(double k)
{
    if (!(k >= 1.0)) return NAN;
    if (k <= 40.0) return logfact_small((int)k);
    const double HALF_LOG_2PI = 0.91893853320467274178;
    const double ik = 1.0 / k, ik2 = ik * ik;
    const double ser = ik * fma(ik2, fma(ik2, fma(ik2, -1.0 / 1680.0, 1.0 / 1260.0), -1.0 / 360.0), 1.0 / 12.0);
    return (fma(k - 0.5, log_pos(k), -k) + HALF_LOG_2PI) + ser;
}

// {sin, cos}(2 pi i/128), the grid of the Box-Muller angle (Philox mode, tm_samplers.cuh).  The kernels that
// sample in their inner loop stage the table in shared memory; everything else reads this copy.
#if defined(__CUDACC__)
__device__ const double kSinCosTabDev[256] = {
#include "tm_sincos_table.inc"
};
#endif
static const double kSinCosTabHost[256] = {
#include "tm_sincos_table.inc"
};
TM_HD const double *sincos_table()
{
#if defined(__CUDA_ARCH__)
    return kSinCosTabDev;
#else
    return kSinCosTabHost;
#endif
}

// sin and cos of 2 pi a/128 for a in [0,128]: i = a rounded to the nearest integer (ties to even, by the
// 1.5*2^52 addition), g = a - i in [-1/2,1/2] (exact), b = (pi/64) g; sin b and cos b - 1 from their Taylor
// series in g^2 (truncation < 4e-18; the three coefficients whose terms stay below 1e-10 carry 21 significant
// bits so that they encode as immediates), then the angle addition with the table entry i mod 128.
// Pure +,-,*,fma and bit operations: the same bits on every machine.
// Full-width FP64 constants of the hot sampler path.  As literals each costs two move instructions per use (a 64-bit
// immediate does not fit an FP64 instruction unless its low word is zero); from the constant bank they are plain operands.
// Same bits either way (the host build, and the oracle's own copy, use the literals).
#if defined(__CUDACC__)
static __constant__ double kMathC[8] = {
    6.93147180559945286227e-01, -1.0 / 6.0, 1.0 / 5.0, 1.0 / 3.0,
    -0x1.4abbce625be53p-16, 0x1.921fb54442d18p-5, 0x1.03c1f081b5ac4p-22, -0x1.3bd3cc9be45dep-10 };
#endif
#if defined(__CUDA_ARCH__)
#define TM_MC(i, lit) kMathC[i]
#else
#define TM_MC(i, lit) (lit)
#endif

TM_HD void sincos_turn128(double a, const double *tab, double &sn, double &cs)
{
    const double magic = 6755399441055744.0; // 1.5 * 2^52
    const double m = a + magic;
    const uint32_t i = (uint32_t)d2u(m) & 127u;
    const double g = a - (m - magic);
    const double z = g * g;
    double sb = fma(z, -0x1.32d2d00000000p-43, 0x1.466bc00000000p-29);
    sb = fma(z, sb, TM_MC(4, -0x1.4abbce625be53p-16));
    sb = fma(z, sb, TM_MC(5, 0x1.921fb54442d18p-5));
    sb = g * sb;                                              // sin b
    double cb = fma(z, -0x1.55d3c00000000p-36, TM_MC(6, 0x1.03c1f081b5ac4p-22));
    cb = fma(z, cb, TM_MC(7, -0x1.3bd3cc9be45dep-10));
    cb = z * cb;                                              // cos b - 1
    const double si = tab[2 * i], ci = tab[2 * i + 1];
    sn = fma(si, cb, fma(ci, sb, si));
    cs = fma(ci, cb, fma(-si, sb, ci));
}

// 1/k for k = 1..64 (IEEE division at compile time: correctly rounded everywhere), index 0 unused:
// the recurrence p_k = p_{k-1} mean / k of the Poisson inversion sampler (tm_samplers.cuh) multiplies instead of dividing.
#define TM_INV_INT_LIST 0.0, 1.0 / 1, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32, 1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37, 1.0 / 38, 1.0 / 39, 1.0 / 40, 1.0 / 41, 1.0 / 42, 1.0 / 43, 1.0 / 44, 1.0 / 45, 1.0 / 46, 1.0 / 47, 1.0 / 48, 1.0 / 49, 1.0 / 50, 1.0 / 51, 1.0 / 52, 1.0 / 53, 1.0 / 54, 1.0 / 55, 1.0 / 56, 1.0 / 57, 1.0 / 58, 1.0 / 59, 1.0 / 60, 1.0 / 61, 1.0 / 62, 1.0 / 63, 1.0 / 64
#if defined(__CUDACC__)
__device__ const double kInvIntDev[65] = { TM_INV_INT_LIST };
#endif
static const double kInvIntHost[65] = { TM_INV_INT_LIST };
TM_HD double inv_int(int k)
{
#if defined(__CUDA_ARCH__)
    return kInvIntDev[k];
#else
    return kInvIntHost[k];
#endif
}

// Philox4x32-10, counter (c0..c3), key (k0,k1).
struct u32x4 { uint32_t x, y, z, w; };

TM_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
#if defined(__CUDA_ARCH__)
    lo = a * b; hi = __umulhi(a, b);
#else
    const uint64_t p = (uint64_t)a * b; lo = (uint32_t)p; hi = (uint32_t)(p >> 32);
#endif
}

TM_HD u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0, l0, h1, l1;
        mulhilo32(0xD2511F53u, c0, h0, l0);
        mulhilo32(0xCD9E8D57u, c2, h1, l1);
        c0 = h1 ^ c1 ^ k0; c1 = l1;
        c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return u32x4{c0, c1, c2, c3};
}

// 52-bit uniform in [0,1) from the top 52 bits of a 64-bit word.
TM_HD double u52(uint32_t lo, uint32_t hi)
{
    const uint64_t w = ((uint64_t)hi << 32) | lo;
    return u2d(0x3FF0000000000000ull | (w >> 12)) - 1.0;
}

} // namespace tml
