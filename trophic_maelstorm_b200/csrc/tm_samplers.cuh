// tm_samplers.cuh -- per-cell uniform streams, the Poisson / Gamma / normal samplers and the
// Dornic cell update.
//
// The sampling LAWS and the order in which uniforms are consumed follow what the reference gets
// from libstdc++ (LS = <bits/random.tcc>, GCC 13):
//   std::normal_distribution  LS:1811-1848   Marsaglia polar, one cached variate (replay mode; a Philox
//                                            stream draws the same law by Box-Muller, no rejection loop)
//   std::poisson_distribution LS:1267-1400   product of uniforms (mean < 12), Devroye rejection (>= 12) (replay mode;
//                                            a Philox stream draws the same law by inversion / PTRS)
//   std::gamma_distribution   LS:2338-2387   Marsaglia-Tsang, u^(1/alpha) boost for alpha < 1
// as called per cell at P/:1055-1059 (DP), P/:1957-1979 (cubic DP), R/:3766-3794 (Rietkerk,
// with the Gaussian shortcuts above 100).  What differs from the reference is the stream: every
// cell owns a counter-based Philox stream (or a slice of host uniforms in replay mode) instead
// of sharing one mt19937_64, so cells are independent and can be sampled in any order.
#pragma once
#include "tm_math.cuh"

namespace tml {

// ---- uniform sources ----------------------------------------------------------------------

// Philox4x32-10: key = seed, counter = (cell, trial, step, block); block b yields uniforms
// 2b (words 0,1) and 2b+1 (words 2,3).
struct PhiloxSrc {
    static constexpr bool kProductLaws = true; // this source draws N(0,1) by Box-Muller and Poisson by inversion / PTRS (see below)
    uint32_t cell, trial, step, k0, k1;
    uint32_t n;       // uniforms handed out so far
    double second;
    TM_HD void begin(uint32_t cell_, uint32_t trial_, uint32_t step_, uint64_t seed)
    {
        cell = cell_; trial = trial_; step = step_;
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); n = 0; second = 0.0;
    }
    TM_HD double next()
    {
        const uint32_t k = n++;
        if (k & 1u) return second;
        const u32x4 o = philox4x32_10(cell, trial, step, k >> 1, k0, k1);
        second = u52(o.z, o.w);
        return u52(o.x, o.y);
    }
    TM_HD void overflowed() {}
    TM_HD uint32_t overflow_count() const { return 0; }
};

// Replay: uniform k of this cell is base[k], k < stride; beyond that 0.75 and a counted overflow.
struct ReplaySrc {
    static constexpr bool kProductLaws = false; // replay follows the reference's libstdc++ algorithms draw for draw
    const double *base;
    uint32_t stride, n, ovf;
    TM_HD void begin(const double *base_, uint32_t stride_) { base = base_; stride = stride_; n = 0; ovf = 0; }
    TM_HD double next()
    {
        const uint32_t k = n++;
        if (k < stride) return base[k];
        ovf++;
        return 0.75;
    }
    TM_HD void overflowed() { ovf++; }
    TM_HD uint32_t overflow_count() const { return ovf; }
};

// ---- samplers -------------------------------------------------------------------------------

struct NormalState { bool have; double saved; };

template <class Src>
TM_HD double normal01(Src &s, NormalState &st)
{
    if (st.have) { st.have = false; return st.saved; }
    if (Src::kProductLaws) {
        // Philox mode: the same N(0,1) law without the rejection loop.  Box-Muller on the next two
        // uniforms: radius from x = 2 - (1+u0 with its last mantissa bit set), in (0,1); angle 2 pi u1.
        const double u0 = s.next(), u1 = s.next();
        const double x = 2.0 - u2d(d2u(1.0 + u0) | 1ull);
        const double r = sqrt(-2.0 * log_pos(x));
        double sn, cs;
        sincos_turn128(128.0 * u1, sincos_table(), sn, cs);
        st.saved = r * sn;
        st.have = true;
        return r * cs;
    }
    double x, y, r2;
    int it = 0;
    do {
        x = 2.0 * s.next() - 1.0;
        y = 2.0 * s.next() - 1.0;
        r2 = fma(x, x, y * y);
    } while ((r2 > 1.0 || r2 == 0.0) && ++it < kMaxReject);
    if (r2 > 1.0 || r2 == 0.0) { s.overflowed(); r2 = 0.5; x = 0.5; y = 0.5; }
    const double mult = sqrt(-2.0 * log_pos(r2) / r2);
    st.saved = x * mult;
    st.have = true;
    return y * mult;
}

// Poisson(mean) for Philox streams: the same law as std::poisson_distribution with algorithms that suit a warp.
//   mean < 12 : inversion by sequential search on ONE uniform (libstdc++ multiplies ~mean+1 uniforms: a Philox block
//               every other iteration, and the warp runs the longest lane's loop);
//   mean >= 12: Hoermann's transformed rejection PTRS (1993): two uniforms per attempt, ~86 % accepted by the
//               squeeze without a logarithm, no set-up beyond one sqrt, one log and two divisions (libstdc++'s
//               Devroye rejection spends ten elementary-function calls on set-up per cell because the reference
//               constructs a fresh distribution object per cell, P/:1055).
template <class Src>
TM_HD long long poisson_product(Src &s, double mean)
{
    if (mean < 12.0) {
        const double u = s.next();
        double p = exp_any(-mean), F = p;
        int k = 0;
        while (u > F && k < 64) { k++; p = p * mean * inv_int(k); F += p; }
        return k;
    }
    const double slam = sqrt(mean), loglam = log_pos(mean);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int it = 0; it < kMaxReject; it++) {
        const double U = s.next() - 0.5;
        const double V = 1.0 - s.next();                       // (0,1]
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + mean + 0.43);
        if (us >= 0.07 && V <= vr) return (long long)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if (log_pos(V) + log_pos(invalpha) - log_pos(a / (us * us) + b) <= -mean + kf * loglam - lgamma_int(kf + 1.0))
            return (long long)kf;
    }
    s.overflowed();
    return (long long)floor(mean);
}

template <class Src>
TM_HD long long poisson_draw(Src &s, double mean)
{
    if (Src::kProductLaws) return poisson_product(s, mean);
    if (mean >= 12.0) {
        NormalState nd{false, 0.0};
        const double m = floor(mean);
        const double lm_thr = log_pos(mean);
        const double lfm = lgamma_int(m + 1.0);
        const double sm = sqrt(m);
        const double pi_4 = 0.7853981633974483096156608458198757;
        const double dxx = sqrt(2.0 * m * log_pos(32.0 * m / pi_4));
        double dm = (m < dxx) ? m : dxx;
        dm = (6.0 < dm) ? dm : 6.0;
        const double d = floor(dm + 0.5);
        const double cx = 2.0 * m + d;
        const double scx = sqrt(cx / 2.0);
        const double i1cx = 1.0 / cx;
        const double c2b = sqrt(pi_4 * cx) * exp_any(i1cx);
        const double cb = 2.0 * cx * exp_any(-d * i1cx * (1.0 + d / 2.0)) / d;
        const double naf = (1.0 - 2.220446049250313e-16) / 2.0;
        const double thr = 2147483647.0 + naf;
        const double spi_2 = 1.2533141373155002512078826424055226;
        const double c1 = sm * spi_2;
        const double c2 = c2b + c1;
        const double c3 = c2 + 1.0;
        const double c4 = c3 + 1.0;
        const double v178 = 0.0128205128205128205128205128205128;
        const double e178 = 1.0129030479320018583185514777512983;
        const double c5 = c4 + e178;
        const double c = cb + c5;
        const double v2cx = 2.0 * (2.0 * m + d);
        double x = 0.0;
        bool reject = true;
        int it = 0;
        do {
            const double u = c * s.next();
            const double e = -log_pos(1.0 - s.next());
            double w = 0.0;
            if (u <= c1) {
                const double n = normal01(s, nd);
                const double y = -fabs(n) * sm - 1.0;
                x = floor(y);
                w = -n * n / 2.0;
                if (x < -m) continue;
            } else if (u <= c2) {
                const double n = normal01(s, nd);
                const double y = 1.0 + fabs(n) * scx;
                x = ceil(y);
                w = y * (2.0 - y) * i1cx;
                if (x > d) continue;
            } else if (u <= c3) {
                x = -1.0;
            } else if (u <= c4) {
                x = 0.0;
            } else if (u <= c5) {
                x = 1.0;
                w = v178;
            } else {
                const double v = -log_pos(1.0 - s.next());
                const double y = d + v * v2cx / d;
                x = ceil(y);
                w = -d * i1cx * (1.0 + y / 2.0);
            }
            reject = (w - e - x * lm_thr > lfm - lgamma_int(x + m + 1.0));
            reject |= (x + m >= thr);
        } while (reject && ++it < kMaxReject);
        if (reject) { s.overflowed(); x = 0.0; }
        return (long long)(x + m + naf);
    }
    long long x = 0;
    double prod = 1.0;
    const double lm_thr = exp_any(-mean);
    do {
        prod *= s.next();
        x += 1;
    } while (prod > lm_thr && x < kMaxProduct);
    return x - 1;
}

template <class Src>
TM_HD double gamma_draw(Src &s, double alpha)
{
    NormalState nd{false, 0.0};
    const double malpha = (alpha < 1.0) ? alpha + 1.0 : alpha;
    const double a1 = malpha - 1.0 / 3.0;
    const double a2 = 1.0 / sqrt(9.0 * a1);
    double u, v, n;
    int it = 0;
    bool again;
    do {
        int it2 = 0;
        do {
            n = normal01(s, nd);
            v = 1.0 + a2 * n;
        } while (v <= 0.0 && ++it2 < kMaxReject);
        if (v <= 0.0) { s.overflowed(); v = 1.0; n = 0.0; }
        v = v * v * v;
        u = s.next();
        again = (u > 1.0 - 0.0331 * n * n * n * n)
             && (log_pos(u) > (0.5 * n * n + a1 * (1.0 - v + log_pos(v))));
    } while (again && ++it < kMaxReject);
    if (again) s.overflowed();
    if (alpha == malpha) return a1 * v * 1.0;
    int it3 = 0;
    do { u = s.next(); } while (u == 0.0 && ++it3 < kMaxReject);
    if (u == 0.0) return 0.0;
    const double ex = (1.0 / alpha) * log_pos(u);
    return exp_any(ex) * a1 * v * 1.0;
}

// ---- the Dornic cell update ------------------------------------------------------------------

struct DornicConsts {
    double diffc;      // D/dx^2                      P/:952  R/:3463
    double lambda;     // 2 beta/(sigma^2 (e^{beta dt}-1))   P/:954  R/:3462
    double ll;         // lambda * lambda_exp (Poisson mean per unit rho)
    double sigma2;     // sigma^2      (DP:  mu = -1 + 2 alpha / sigma^2,   P/:1054)
    double sigma2_1;   // 1/sigma^2    (Rietkerk: mu = -1 + 2 alpha * sigma2_1, R/:3766)
};

enum { kDP = 0, kDPCubic = 1, kRietkerk = 2 };

// rho* for one live cell (the caller has already applied the alpha==0 && rho==0 skip).
template <int MODEL, class Src>
TM_HD double dornic_cell(const DornicConsts &k, Src &s, double alpha_i, double rho_i)
{
    const double ziggy = k.ll * rho_i;
    if (MODEL == kDP) {
        const double mu = -1.0 + 2.0 * alpha_i / k.sigma2;
        const double gru = mu + 1.0 + (double)poisson_draw(s, ziggy); // mean 0 still burns one uniform
        return gamma_draw(s, gru) / k.lambda;
    } else if (MODEL == kDPCubic) {
        const double mu = -1.0 + 2.0 * alpha_i / k.sigma2;
        double gru = mu + 1.0;
        if (ziggy != 0.0) gru = mu + 1.0 + (double)poisson_draw(s, ziggy);
        return gamma_draw(s, gru) / k.lambda;
    } else {
        NormalState norm{false, 0.0}; // the reference's explicit `norm` (R/:3631), per cell here
        const double mu = -1.0 + 2.0 * alpha_i * k.sigma2_1;
        double gru;
        if (ziggy == 0.0) {
            gru = mu + 1.0;
        } else if (ziggy > 100) {
            const long long gauss = (long long)fma(normal01(s, norm), sqrt(ziggy), ziggy);
            gru = mu + 1.0 + (double)gauss;
        } else {
            gru = mu + 1.0 + (double)poisson_draw(s, ziggy);
        }
        if (gru > 100) {
            const double gauss = fma(normal01(s, norm), sqrt(gru), gru);
            return gauss / k.lambda;
        }
        return gamma_draw(s, gru) / k.lambda;
    }
}

} // namespace tml
