/*
 * oracle/tm_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT (see tm_oracle.h).
 *
 * Plain C restatement of the reference's Dornic lattice step.  Compile with
 * -ffp-contract=off (oracle/Makefile does): every fused multiply-add below is an explicit
 * fma(), everything else rounds once per operation, so the result is a pure function of the
 * inputs on any IEEE-754 machine.
 *
 * Citation shorthands are defined in tm_oracle.h (P/, R/, F/, LS).
 */
#include "tm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ================================================================================== */
/* 1. Elementary functions of the sampler spec (DESIGN.md "Sampler spec").           */
/*    Only +,-,*,/,sqrt,fma and integer bit operations: bit-reproducible on CPU/GPU.  */
/* ================================================================================== */

static inline uint64_t d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double   u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* Natural log for finite x > 0 (subnormals excluded by construction of the callers).
 * Argument reduction x = 2^k m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = f/(2+f), f = m-1,
 * with the classical degree-14 odd minimax polynomial (Sun fdlibm e_log.c coefficients).
 * Replaces std::log at LS:1832 (polar normal), LS:1344/1382 (-log(1-u)), LS:2372 (gamma test). */
double tmo_log(double x)
{
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    static const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                        Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                        Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                        Lg7 = 1.479819860511658591e-01;
    if (!(x > 0.0)) return (x == 0.0) ? -INFINITY : NAN;
    if (x == INFINITY) return x;
    uint64_t b = d2u(x);
    int k = (int)(b >> 52) - 1023;
    uint64_t mant = b & 0x000FFFFFFFFFFFFFull;
    if (mant > 0x6A09E667F3BCCull) { /* m > sqrt(2): halve it */
        b = mant | 0x3FE0000000000000ull; k += 1;
    } else {
        b = mant | 0x3FF0000000000000ull;
    }
    double m = u2d(b);
    double f = m - 1.0;
    double dk = (double)k;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
    double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    return dk * ln2_hi - ((hfsq - fma(s, hfsq + R, dk * ln2_lo)) - f);
}

/* exp(x): k = rint(x log2 e), r = x - k ln2 (two-part), degree-13 Taylor in r, scaled by 2^k in
 * two exact-then-rounded steps.  Replaces std::exp at LS:1295 (Poisson threshold e^-mean),
 * LS:1288-1290 (Devroye constants) and the pow() of LS:2384 (via exp(p log u)). */
double tmo_exp(double x)
{
    static const double LOG2E = 1.44269504088896338700e+00;
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.2) return 0.0;
    double kf = rint(x * LOG2E);
    double r = fma(kf, -ln2_hi, x);
    r = fma(kf, -ln2_lo, r);
    double q = 1.0 / 6227020800.0;          /* 1/13! */
    q = fma(q, r, 1.0 / 479001600.0);       /* 1/12! */
    q = fma(q, r, 1.0 / 39916800.0);
    q = fma(q, r, 1.0 / 3628800.0);
    q = fma(q, r, 1.0 / 362880.0);
    q = fma(q, r, 1.0 / 40320.0);
    q = fma(q, r, 1.0 / 5040.0);
    q = fma(q, r, 1.0 / 720.0);
    q = fma(q, r, 1.0 / 120.0);
    q = fma(q, r, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    double y = 1.0 + fma(r * r, q, r);
    int k = (int)kf;
    int k1 = k / 2, k2 = k - k1;
    y *= u2d((uint64_t)(k1 + 1023) << 52);
    y *= u2d((uint64_t)(k2 + 1023) << 52);
    return y;
}

/* log Gamma(k) for integer-valued k >= 1 (the only arguments LS:1280 and LS:1388 ever pass:
 * floor(mean)+1 and x+m+1 with x integer-valued).  Exact table of log((k-1)!) rounded to double
 * for k <= 40, Stirling series beyond. */
static const double TMO_LOGFACT[40] = {
#include "tm_logfact_table.inc"
};

double tmo_lgamma_int(double k)
{
    if (!(k >= 1.0)) return NAN;
    if (k <= 40.0) return TMO_LOGFACT[(int)k - 1];
    static const double HALF_LOG_2PI = 0.91893853320467274178;
    double ik = 1.0 / k, ik2 = ik * ik;
    double ser = ik * fma(ik2, fma(ik2, fma(ik2, -1.0 / 1680.0, 1.0 / 1260.0), -1.0 / 360.0), 1.0 / 12.0);
    return (fma(k - 0.5, tmo_log(k), -k) + HALF_LOG_2PI) + ser;
}

/* sin and cos of 2 pi a/128, a in [0,128] -- the Box-Muller angle of the product spec (DESIGN.md
 * section 3): nearest grid point i (ties to even through the 1.5*2^52 addition), g = a - i in
 * [-1/2,1/2] exactly, b = (pi/64) g; sin b and cos b - 1 by Taylor series in g^2 with the
 * coefficients (+-)(pi/64)^n / n! (S2, S3, C3 rounded to 21 significant bits, as the spec says), then
 * the angle addition with {sin, cos}(2 pi i/128) from the correctly rounded table of gen_tables.py. */
static const double tmo_sincos_tab[256] = {
#include "tm_sincos_table.inc"
};
void tmo_sincos_turn128(double a, double *sn, double *cs)
{
    static const double S[4] = { 0x1.921fb54442d18p-5, -0x1.4abbce625be53p-16, 0x1.466bc00000000p-29, -0x1.32d2d00000000p-43 };
    static const double C[4] = { 0.0, -0x1.3bd3cc9be45dep-10, 0x1.03c1f081b5ac4p-22, -0x1.55d3c00000000p-36 };
    const double magic = 6755399441055744.0;
    volatile double m = a + magic;          /* volatile: the rounding addition must really happen */
    uint32_t i = (uint32_t)d2u(m) & 127u;
    double g = a - (m - magic);
    double z = g * g;
    double sb = g * fma(z, fma(z, fma(z, S[3], S[2]), S[1]), S[0]);
    double cb = z * fma(z, fma(z, C[3], C[2]), C[1]);
    double si = tmo_sincos_tab[2 * i], ci = tmo_sincos_tab[2 * i + 1];
    *sn = fma(si, cb, fma(ci, sb, si));
    *cs = fma(ci, cb, fma(-si, sb, ci));
}

/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator that replaces
 * std::mt19937_64 (P/:975, R/:3513) in the product. */
void tmo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 52-bit uniform in [0,1): top 52 bits of w as the mantissa of a double in [1,2), minus 1. */
double tmo_u52_from_bits(uint64_t w)
{
    return u2d(0x3FF0000000000000ull | (w >> 12)) - 1.0;
}

/* ================================================================================== */
/* 2. Uniform streams                                                                 */
/* ================================================================================== */

typedef struct {
    tmo_rng *g;
    uint64_t cell;     /* REPLAY / PHILOX */
    uint64_t k;        /* draw index within the cell */
    double second;     /* PHILOX: second uniform of the current block */
} cstream;

static void cs_begin_cell(cstream *s, tmo_rng *g, uint64_t cell)
{
    s->g = g; s->cell = cell; s->k = 0; s->second = 0.0;
}

static double next_u(cstream *s)
{
    tmo_rng *g = s->g;
    g->draws++;
    switch (g->kind) {
    case TMO_SRC_SEQ:
        if (g->pos < g->n_u) return g->u[g->pos++];
        g->overflow++; return 0.75;
    case TMO_SRC_REPLAY: {
        uint64_t k = s->k++;
        if (k < g->stride && s->cell * g->stride + k < g->n_u) return g->u[s->cell * g->stride + k];
        g->overflow++; return 0.75;
    }
    default: { /* TMO_SRC_PHILOX: block b = k/2 yields uniforms 2b (words 0,1) and 2b+1 (words 2,3) */
        uint64_t k = s->k++;
        if (k & 1) return s->second;
        uint32_t ctr[4] = { (uint32_t)s->cell, g->trial, g->step, (uint32_t)(k >> 1) };
        uint32_t key[2] = { (uint32_t)g->seed, (uint32_t)(g->seed >> 32) };
        uint32_t o[4];
        tmo_philox4x32_10(ctr, key, o);
        s->second = tmo_u52_from_bits(((uint64_t)o[3] << 32) | o[2]);
        return tmo_u52_from_bits(((uint64_t)o[1] << 32) | o[0]);
    }
    }
}

/* ================================================================================== */
/* 3. Samplers: restatement of libstdc++ 13 (LS)                                      */
/* ================================================================================== */

typedef struct { int saved_available; double saved; } normal_state;

/* Spec addition (not in LS): every rejection loop is bounded, so that a replay buffer that
 * runs dry -- or a GPU thread fed garbage -- cannot spin forever.  The bounds are never reached
 * with genuine uniforms (polar: 0.215^64; Devroye/Marsaglia-Tsang accept with p > 0.5). */
#define TM_MAX_REJECT 64
#define TM_MAX_PRODUCT 4096

/* std::normal_distribution<double>(0,1)::operator(), Marsaglia polar with one cached
 * variate: LS:1811-1848. */
static double normal01(cstream *s, normal_state *st)
{
    if (st->saved_available) {
        st->saved_available = 0;
        return st->saved;
    }
    if (s->g->kind == TMO_SRC_PHILOX) {
        /* Product spec, Philox streams only: the N(0,1) law of LS:1811-1848 drawn without its
         * rejection loop -- Box-Muller on the next two uniforms.  Radius from x = 2 - m, m = 1+u0
         * with its last mantissa bit set (x in (0,1), never 0 or 1); angle 2 pi u1.  The sequential
         * and replay sources above keep the reference's polar loop draw for draw. */
        double u0 = next_u(s), u1 = next_u(s);
        double xr = 2.0 - u2d(d2u(1.0 + u0) | 1ull);
        double r = sqrt(-2.0 * tmo_log(xr));
        double sn, cs;
        tmo_sincos_turn128(128.0 * u1, &sn, &cs);
        st->saved = r * sn;
        st->saved_available = 1;
        return r * cs;
    }
    double x, y, r2;
    int it = 0;
    do {
        x = 2.0 * next_u(s) - 1.0;
        y = 2.0 * next_u(s) - 1.0;
        r2 = fma(x, x, y * y);
    } while ((r2 > 1.0 || r2 == 0.0) && ++it < TM_MAX_REJECT);
    if (r2 > 1.0 || r2 == 0.0) { s->g->overflow++; r2 = 0.5; x = 0.5; y = 0.5; }
    double mult = sqrt(-2.0 * tmo_log(r2) / r2);
    st->saved = x * mult;
    st->saved_available = 1;
    return y * mult;
}


/* std::poisson_distribution<int>(mean)::operator() for a freshly constructed object
 * (the reference constructs one per cell: P/:1055, R/:3781).  Parameter set-up LS:1267-1296,
 * draw LS:1310-1400: Devroye's rejection for mean >= 12, product of uniforms below. */
/* Product spec, Philox streams only: Poisson(mean) by inversion on one uniform (mean < 12) or by
 * Hoermann's PTRS transformed rejection (mean >= 12) -- same law as LS, algorithms that suit a warp
 * (DESIGN.md section 3).  Own restatement; 1/k comes from an IEEE division, which is what the
 * product's compile-time table holds. */
static long poisson_product(cstream *s, double mean)
{
    if (mean < 12.0) {
        double u = next_u(s);
        double p = tmo_exp(-mean), F = p;
        int k = 0;
        while (u > F && k < 64) { k++; p = p * mean * (1.0 / (double)k); F += p; }
        return k;
    }
    double slam = sqrt(mean), loglam = tmo_log(mean);
    double b = 0.931 + 2.53 * slam;
    double a = -0.059 + 0.02483 * b;
    double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int it = 0; it < TM_MAX_REJECT; it++) {
        double U = next_u(s) - 0.5;
        double V = 1.0 - next_u(s);
        double us = 0.5 - fabs(U);
        double kf = floor((2.0 * a / us + b) * U + mean + 0.43);
        if (us >= 0.07 && V <= vr) return (long)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if (tmo_log(V) + tmo_log(invalpha) - tmo_log(a / (us * us) + b) <= -mean + kf * loglam - tmo_lgamma_int(kf + 1.0))
            return (long)kf;
    }
    s->g->overflow++;
    return (long)floor(mean);
}

static long poisson_draw(cstream *s, double mean)
{
    if (s->g->kind == TMO_SRC_PHILOX) return poisson_product(s, mean);
    if (mean >= 12.0) {
        normal_state nd = {0, 0.0}; /* the distribution's own _M_nd, fresh (LS: poisson ctor) */
        /* _M_initialize, LS:1276-1291 */
        const double m = floor(mean);
        const double lm_thr = tmo_log(mean);
        const double lfm = tmo_lgamma_int(m + 1.0);
        const double sm = sqrt(m);
        const double pi_4 = 0.7853981633974483096156608458198757;
        const double dxx = sqrt(2.0 * m * tmo_log(32.0 * m / pi_4));
        double dm = (m < dxx) ? m : dxx;          /* std::min(m, dx) */
        dm = (6.0 < dm) ? dm : 6.0;               /* std::max(6.0, .) */
        const double d = floor(dm + 0.5);         /* std::round, argument >= 6 */
        const double cx = 2.0 * m + d;
        const double scx = sqrt(cx / 2.0);
        const double i1cx = 1.0 / cx;
        const double c2b = sqrt(pi_4 * cx) * tmo_exp(i1cx);
        const double cb = 2.0 * cx * tmo_exp(-d * i1cx * (1.0 + d / 2.0)) / d;
        /* operator(), LS:1322-1398 */
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
        int reject = 1, it = 0;
        do {
            const double u = c * next_u(s);
            const double e = -tmo_log(1.0 - next_u(s));
            double w = 0.0;
            if (u <= c1) {
                const double n = normal01(s, &nd);
                const double y = -fabs(n) * sm - 1.0;
                x = floor(y);
                w = -n * n / 2.0;
                if (x < -m) continue;
            } else if (u <= c2) {
                const double n = normal01(s, &nd);
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
                const double v = -tmo_log(1.0 - next_u(s));
                const double y = d + v * v2cx / d;
                x = ceil(y);
                w = -d * i1cx * (1.0 + y / 2.0);
            }
            reject = (w - e - x * lm_thr > lfm - tmo_lgamma_int(x + m + 1.0));
            reject |= (x + m >= thr);
        } while (reject && ++it < TM_MAX_REJECT);
        if (reject) { s->g->overflow++; x = 0.0; }
        return (long)(x + m + naf);
    } else {
        long x = 0;
        double prod = 1.0;
        const double lm_thr = tmo_exp(-mean); /* LS:1295 */
        do {
            prod *= next_u(s);
            x += 1;
        } while (prod > lm_thr && x < TM_MAX_PRODUCT);
        return x - 1;
    }
}

/* std::gamma_distribution<double>(alpha, 1.0)::operator() for a freshly constructed object
 * (P/:1057, R/:3792): Marsaglia-Tsang with the alpha < 1 boost; set-up LS:2338-2345, draw
 * LS:2352-2387.  std::pow(u, 1/alpha) is restated as exp(log(u)/alpha). */
static double gamma_draw(cstream *s, double alpha)
{
    normal_state nd = {0, 0.0};
    const double malpha = (alpha < 1.0) ? alpha + 1.0 : alpha;
    const double a1 = malpha - 1.0 / 3.0;
    const double a2 = 1.0 / sqrt(9.0 * a1);
    double u, v, n;
    int it = 0, again;
    do {
        int it2 = 0;
        do {
            n = normal01(s, &nd);
            v = 1.0 + a2 * n;
        } while (v <= 0.0 && ++it2 < TM_MAX_REJECT);
        if (v <= 0.0) { s->g->overflow++; v = 1.0; n = 0.0; }
        v = v * v * v;
        u = next_u(s);
        again = (u > 1.0 - 0.0331 * n * n * n * n)
             && (tmo_log(u) > (0.5 * n * n + a1 * (1.0 - v + tmo_log(v))));
    } while (again && ++it < TM_MAX_REJECT);
    if (again) s->g->overflow++;
    if (alpha == malpha)
        return a1 * v * 1.0;
    int it3 = 0;
    do {
        u = next_u(s);
    } while (u == 0.0 && ++it3 < TM_MAX_REJECT);
    if (u == 0.0) return 0.0;
    /* pow(u, 1/alpha): u in (0,1); alpha == 0 gives u^inf = 0 exactly as LS does */
    const double ex = (1.0 / alpha) * tmo_log(u);
    return tmo_exp(ex) * a1 * v * 1.0;
}

/* test hooks: drive the samplers from a flat array of uniforms (SEQ semantics) */
static void seq_rng(tmo_rng *g, const double *u, uint64_t n, uint64_t pos)
{
    memset(g, 0, sizeof *g);
    g->kind = TMO_SRC_SEQ; g->u = u; g->n_u = n; g->pos = pos;
}
double tmo_test_normal_pair(const double *u, uint64_t n, uint64_t *pos, double *second)
{
    tmo_rng g; seq_rng(&g, u, n, *pos);
    cstream s; cs_begin_cell(&s, &g, 0);
    normal_state st = {0, 0.0};
    double a = normal01(&s, &st);
    *second = normal01(&s, &st);
    *pos = g.pos;
    return a;
}
/* the first two normal variates of the Philox stream of one cell (Box-Muller law of the product spec) */
double tmo_test_normal_pair_philox(uint64_t seed, uint32_t cell, uint32_t trial, uint32_t step, double *second)
{
    tmo_rng g; memset(&g, 0, sizeof g);
    g.kind = TMO_SRC_PHILOX; g.seed = seed; g.trial = trial; g.step = step;
    cstream s; cs_begin_cell(&s, &g, cell);
    normal_state st = {0, 0.0};
    double a = normal01(&s, &st);
    *second = normal01(&s, &st);
    return a;
}
long tmo_test_poisson_philox(uint64_t seed, uint32_t cell, uint32_t trial, uint32_t step, double mean)
{
    tmo_rng g; memset(&g, 0, sizeof g);
    g.kind = TMO_SRC_PHILOX; g.seed = seed; g.trial = trial; g.step = step;
    cstream s; cs_begin_cell(&s, &g, cell);
    return poisson_draw(&s, mean);
}
long tmo_test_poisson(const double *u, uint64_t n, uint64_t *pos, double mean)
{
    tmo_rng g; seq_rng(&g, u, n, *pos);
    cstream s; cs_begin_cell(&s, &g, 0);
    long r = poisson_draw(&s, mean);
    *pos = g.pos;
    return r;
}
double tmo_test_gamma(const double *u, uint64_t n, uint64_t *pos, double shape)
{
    tmo_rng g; seq_rng(&g, u, n, *pos);
    cstream s; cs_begin_cell(&s, &g, 0);
    double r = gamma_draw(&s, shape);
    *pos = g.pos;
    return r;
}

/* ================================================================================== */
/* 4. The Dornic cell update                                             */
/* ================================================================================== */

typedef struct {
    double diffc;       /* D/dx^2            P/:952   R/:3463 */
    double beta;        /* P/:953  R/:3461 */
    double lambda_exp;  /* exp(beta dt)      P/:954   R/:3462 */
    double lambda;      /* 2 beta/(sigma^2 (lambda_exp-1)) */
    double sigma2;      /* sigma*sigma */
    double sigma2_1;    /* 1/(sigma*sigma)   R/:3463 */
} dornic_consts;

static void derive_consts(const tmo_params *p, dornic_consts *k)
{
    const double dx2 = p->dx * p->dx;
    if (p->model == TMO_MODEL_DP || p->model == TMO_MODEL_DP_CUBIC) {
        k->diffc = p->D / (p->dx * p->dx);
        k->beta = p->a - 4 * p->D / (p->dx * p->dx);
        k->lambda_exp = exp(k->beta * p->dt);
        k->lambda = 2 * k->beta / (p->sigma * p->sigma * (k->lambda_exp - 1.0));
        k->sigma2 = p->sigma * p->sigma;
        k->sigma2_1 = 1 / k->sigma2;
    } else {
        k->beta = -p->d - 4 * p->D0 / dx2;
        k->lambda_exp = exp(k->beta * p->dt);
        k->lambda = 2 * k->beta / (p->sigma0 * p->sigma0 * (k->lambda_exp - 1.0));
        k->diffc = p->D0 / dx2;
        k->sigma2 = p->sigma0 * p->sigma0;
        k->sigma2_1 = 1 / (p->sigma0 * p->sigma0);
    }
}

/* rho* for one live cell.  alpha_i already computed; rho_i = Rho_dt[i] (== DRho[i]). */
static double dornic_cell(const tmo_params *p, const dornic_consts *k, cstream *s,
                          normal_state *norm, double alpha_i, double rho_i)
{
    double mu, ziggy, gru;
    switch (p->model) {
    case TMO_MODEL_DP: /* P/:1054-1059 */
        mu = -1.0 + 2.0 * alpha_i / k->sigma2;
        ziggy = k->lambda * k->lambda_exp * rho_i;
        /* poisson_distribution<int>(0.0) still burns one uniform (LS:1390-1397) */
        gru = mu + 1.0 + (double)poisson_draw(s, ziggy);
        return gamma_draw(s, gru) / k->lambda;
    case TMO_MODEL_DP_CUBIC: /* P/:1957-1977 */
        mu = -1.0 + 2.0 * alpha_i / k->sigma2;
        ziggy = k->lambda * k->lambda_exp * rho_i;
        if (ziggy == 0.0) gru = mu + 1.0;
        else              gru = mu + 1.0 + (double)poisson_draw(s, ziggy);
        return gamma_draw(s, gru) / k->lambda;
    default: /* Rietkerk vegetation, R/:3766-3794 (= F/:4168-4196) */
        mu = -1.0 + 2.0 * alpha_i * k->sigma2_1;
        ziggy = k->lambda * k->lambda_exp * rho_i;
        if (ziggy == 0.0) {
            gru = mu + 1.0;
        } else if (ziggy > 100) {
            long gauss = (long)fma(normal01(s, norm), sqrt(ziggy), ziggy);
            gru = mu + 1.0 + (double)gauss;
        } else {
            gru = mu + 1.0 + (double)poisson_draw(s, ziggy);
        }
        if (gru > 100) {
            double gauss = fma(normal01(s, norm), sqrt(gru), gru);
            return gauss / k->lambda;
        }
        return gamma_draw(s, gru) / k->lambda;
    }
}

/* The cell loop: P/:1032-1061 (S8 table, P/:416-436), R/:3752-3809 (Sq4 table, R/:1609-1628). */
static void dornic_sweep(const tmo_params *p, const dornic_consts *k, tmo_rng *rng,
                         const double *prev, double *out)
{
    const int L = p->L;
    cstream s;
    normal_state norm_cell;
    normal_state norm_seq = { rng->norm_saved_available, rng->norm_saved };
    cs_begin_cell(&s, rng, 0);
    for (int r = 0; r < L; r++) {
        const int up = ((r - 1) % L + L) % L, dn = ((r + 1) % L + L) % L;
        for (int q = 0; q < L; q++) {
            const int lf = ((q - 1) % L + L) % L, rt = ((q + 1) % L + L) % L;
            const int i = r * L + q;
            const double alpha_i = k->diffc * (prev[up * L + q] + prev[dn * L + q]
                                             + prev[r * L + lf] + prev[r * L + rt]);
            if (alpha_i == 0 && prev[i] == 0) { out[i] = prev[i]; continue; }
            normal_state *norm = &norm_seq;
            if (rng->kind != TMO_SRC_SEQ) {
                cs_begin_cell(&s, rng, (uint64_t)i);
                norm_cell.saved_available = 0; norm_cell.saved = 0.0;
                norm = &norm_cell;
            }
            out[i] = dornic_cell(p, k, &s, norm, alpha_i, prev[i]);
        }
    }
    rng->norm_saved_available = norm_seq.saved_available;
    rng->norm_saved = norm_seq.saved;
}

/* The grazer's Dornic update with upwind advection, R/:3908-4114 (SPB >= 2, species s = 1), for every cell in the
 * reference's order.  prev = DRho[s] (state of the previous step); gamma[s] is clamped at 1 in place (R/:3915);
 * veff receives the sampled advection vector {first, second} of every cell (R/:3993-3996: drawn for EVERY cell, before
 * the skip test); out = rho* (cells skipped by R/:4048 keep prev).  trig_libm != 0: cos/sin of libm on 2*PI*ran as the
 * reference computes them (sequential mode, pins the restatement); 0: the product's bit-reproducible sincos_turn128. */
int tmo_grazer_dornic(const tmo_graz_params *p, tmo_rng *rng, const double *prev, double *gamma, double *out, double *veff,
                      int trig_libm)
{
    if (!p || p->L < 1) return -1;
    const int L = p->L;
    const double dx = p->dx, dt = p->dt;
    const double dx2 = dx * dx, dx1 = 1 / dx, dx1_2 = 1 / (dx2), dxb2 = dx / 2.0;   /* R/:3455 */
    const double sigma2_1 = 1 / (p->sigma * p->sigma);                              /* R/:3463 */
    double pi = 0.0, p16 = 1.0;                                                     /* the reference's PI, header :241-284 */
    for (int n = 0; n <= 14; n++) {
        pi = (1.0 / p16 * (4.0 / (8 * n + 1.0) - 2.0 / (8 * n + 4.0) - 1.0 / (8 * n + 5.0) - 1.0 / (8 * n + 6.0))) + pi;
        p16 *= 16.0;
    }
    cstream s;
    normal_state norm_cell;
    normal_state norm_seq = { rng->norm_saved_available, rng->norm_saved };
    cs_begin_cell(&s, rng, 0);
    for (int r = 0; r < L; r++) {
        const int up = ((r - 1) % L + L) % L, dn = ((r + 1) % L + L) % L;
        for (int q = 0; q < L; q++) {
            const int lf = ((q - 1) % L + L) % L, rt = ((q + 1) % L + L) % L;
            const int i = r * L + q;
            if (gamma[i] > 1) gamma[i] = 1;                                          /* R/:3915 */
            normal_state *norm = &norm_seq;
            if (rng->kind != TMO_SRC_SEQ) {
                cs_begin_cell(&s, rng, (uint64_t)i);
                norm_cell.saved_available = 0; norm_cell.saved = 0.0;
                norm = &norm_cell;
            }
            const double ran = next_u(&s);                                           /* R/:3993 */
            double vx, vy;
            if (trig_libm) { vx = p->v * cos(2 * pi * ran); vy = p->v * sin(2 * pi * ran); }   /* R/:3995 */
            else { double sn, cs; tmo_sincos_turn128(128.0 * ran, &sn, &cs); vx = p->v * cs; vy = p->v * sn; }
            veff[2 * i] = vx; veff[2 * i + 1] = vy;
            const double vx_abs = vx * ((0 < vx) - (vx < 0)), vy_abs = vy * ((0 < vy) - (vy < 0));   /* R/:4000, sgn R/:25 */
            const double gp = (1 - gamma[i]);
            double dex = (p->D - (dxb2) * (gp * vx_abs * (1 - dt * dx1 * gp * vx_abs))) * dx1_2;    /* R/:4008 */
            double dey = (p->D - (dxb2) * (gp * vy_abs * (1 - dt * dx1 * gp * vy_abs))) * dx1_2;    /* R/:4009 */
            if (dex < p->D0floor) dex = p->D0floor;                                  /* R/:4014-4018 */
            if (dey < p->D0floor) dey = p->D0floor;
            const double hx = prev[r * L + ((0 < vx) ? lf : rt)];                    /* nR2[i][1][sgn_index(vx)], R/:30-35 */
            const double hy = prev[((0 < vy) ? up : dn) * L + q];                    /* nR2[i][0][sgn_index(vy)] */
            const double alpha_i = gamma[i] * (dey * (prev[up * L + q] + prev[dn * L + q]) + dex * (prev[r * L + lf] + prev[r * L + rt]))
                                 + dx1 * (gp) * (vx_abs * hx + vy_abs * hy);         /* R/:4044-4046 */
            if (alpha_i == 0 && prev[i] == 0) { out[i] = prev[i]; continue; }        /* R/:4048 */
            const double beta = -p->M - 2 * gamma[i] * (dex + dey) - (1 - gamma[i]) * (vx_abs + vy_abs) * dx1;   /* R/:4057 */
            dornic_consts k;
            k.beta = beta; k.lambda_exp = trig_libm ? exp((beta) * dt) : tmo_exp((beta) * dt);   /* R/:4059 */
            k.lambda = 2 * (beta) * (sigma2_1) / (k.lambda_exp - 1.0);
            k.sigma2 = p->sigma * p->sigma; k.sigma2_1 = sigma2_1; k.diffc = 0.0;
            tmo_params pm; memset(&pm, 0, sizeof pm); pm.model = TMO_MODEL_RIETKERK;  /* the same sampling ladder as R/:3766-3794 */
            out[i] = dornic_cell(&pm, &k, &s, norm, alpha_i, prev[i]);               /* R/:4061-4090 */
        }
    }
    rng->norm_saved_available = norm_seq.saved_available;
    rng->norm_saved = norm_seq.saved;
    return 0;
}

/* ================================================================================== */
/* 5. Deterministic remainders                                                        */
/* ================================================================================== */

/* f_2Dor, R/:2359-2386.  s = {P,W,O} state, f = RHS; neighbour order up,down,left,right. */
static void f_2Dor(const tmo_params *p, const double *S, double *f)
{
    const int L = p->L; const int64_t n = (int64_t)L * L;
    const double *P = S, *W = S + n, *O = S + 2 * n;
    const double dx2 = p->dx * p->dx;
    const double Dxd2_1 = p->D1 / dx2, Dxd2_2 = p->D2 / dx2; /* R/:3479-3480 */
    const double cgmax = p->rc * p->gmax;
    for (int r = 0; r < L; r++) {
        const int up = ((r - 1) % L + L) % L, dn = ((r + 1) % L + L) % L;
        for (int q = 0; q < L; q++) {
            const int lf = ((q - 1) % L + L) % L, rt = ((q + 1) % L + L) % L;
            const int64_t i = (int64_t)r * L + q, iu = (int64_t)up * L + q, id = (int64_t)dn * L + q,
                          il = (int64_t)r * L + lf, ir = (int64_t)r * L + rt;
            f[i] = cgmax * W[i] * P[i] / (W[i] + p->K1);
            f[n + i] = p->alpha * (P[i] + p->K2 * p->W0) / (P[i] + p->K2) * O[i]
                     - p->gmax * W[i] * P[i] / (W[i] + p->K1) - p->rW * W[i]
                     + Dxd2_1 * (W[iu] + W[id] + W[il] + W[ir] - 4 * W[i]);
            f[2 * n + i] = p->a - p->alpha * (P[i] + p->K2 * p->W0) / (P[i] + p->K2) * O[i]
                     + Dxd2_2 * (O[iu] + O[id] + O[il] + O[ir] - 4 * O[i]);
        }
    }
}

/* f_2D, R/:1718-1743: the full right-hand side (noise-off path). */
static void f_2D(const tmo_params *p, const double *S, double *f)
{
    const int L = p->L; const int64_t n = (int64_t)L * L;
    const double *P = S, *W = S + n, *O = S + 2 * n;
    const double dx2 = p->dx * p->dx;
    const double Dxd2_0 = p->D0 / dx2, Dxd2_1 = p->D1 / dx2, Dxd2_2 = p->D2 / dx2;
    for (int r = 0; r < L; r++) {
        const int up = ((r - 1) % L + L) % L, dn = ((r + 1) % L + L) % L;
        for (int q = 0; q < L; q++) {
            const int lf = ((q - 1) % L + L) % L, rt = ((q + 1) % L + L) % L;
            const int64_t i = (int64_t)r * L + q, iu = (int64_t)up * L + q, id = (int64_t)dn * L + q,
                          il = (int64_t)r * L + lf, ir = (int64_t)r * L + rt;
            f[i] = p->rc * p->gmax * W[i] * P[i] / (W[i] + p->K1) - p->d * P[i]
                 + Dxd2_0 * (P[iu] + P[id] + P[il] + P[ir] - 4 * P[i]);
            f[n + i] = p->alpha * (P[i] + p->K2 * p->W0) / (P[i] + p->K2) * O[i]
                     - p->gmax * W[i] * P[i] / (W[i] + p->K1) - p->rW * W[i]
                     + Dxd2_1 * (W[iu] + W[id] + W[il] + W[ir] - 4 * W[i]);
            f[2 * n + i] = p->a - p->alpha * (P[i] + p->K2 * p->W0) / (P[i] + p->K2) * O[i]
                     + Dxd2_2 * (O[iu] + O[id] + O[il] + O[ir] - 4 * O[i]);
        }
    }
}

/* Classical RK4 over the three fields: RK4_Integrate_Stochastic_1Sp R/:2389-2440 (f_2Dor)
 * and RK4_Integrate R/:1753-1814 (f_2D); dt2 = dt/2, dt6 = dt/6 (R/:170). */
static int rk4_3field(const tmo_params *p, double *S, int full_rhs)
{
    const int64_t n3 = 3 * (int64_t)p->L * p->L;
    double *buf = (double *)malloc(sizeof(double) * 5 * n3);
    if (!buf) return -1;
    double *tsar = buf, *K1 = buf + n3, *K2 = buf + 2 * n3, *K3 = buf + 3 * n3, *K4 = buf + 4 * n3;
    const double dt = p->dt, dt2 = dt / 2.0, dt6 = dt / 6.0;
    void (*f)(const tmo_params *, const double *, double *) = full_rhs ? f_2D : f_2Dor;
    f(p, S, K1);
    for (int64_t i = 0; i < n3; i++) tsar[i] = S[i] + dt2 * K1[i];
    f(p, tsar, K2);
    for (int64_t i = 0; i < n3; i++) tsar[i] = S[i] + dt2 * K2[i];
    f(p, tsar, K3);
    for (int64_t i = 0; i < n3; i++) tsar[i] = S[i] + dt * K3[i];
    f(p, tsar, K4);
    for (int64_t i = 0; i < n3; i++) S[i] += dt6 * (K1[i] + 2.0 * K2[i] + 2.0 * K3[i] + K4[i]);
    free(buf);
    return 0;
}

/* ================================================================================== */
/* 6. One full time step                                                              */
/* ================================================================================== */

/* RK4 alone on a 3-field state: full_rhs = 0 -> RK4_Integrate_Stochastic_1Sp / f_2Dor (R/:2389),
 * 1 -> RK4_Integrate / f_2D (R/:1753). */
int tmo_rk4(const tmo_params *p, double *f, int full_rhs)
{
    if (!p || !f || p->L < 1) return -1;
    return rk4_3field(p, f, full_rhs);
}

int tmo_dornic_only(const tmo_params *p, tmo_rng *rng, const double *prev, double *out)
{
    if (!p || !rng || !prev || !out || p->L < 1 || p->model == TMO_MODEL_RIETKERK_DET) return -1;
    dornic_consts k; derive_consts(p, &k);
    dornic_sweep(p, &k, rng, prev, out);
    return 0;
}

int tmo_step(const tmo_params *p, tmo_rng *rng, double *f)
{
    if (!p || !f || p->L < 1) return -1;
    const int64_t n = (int64_t)p->L * p->L;
    dornic_consts k;
    if (p->model == TMO_MODEL_RIETKERK_DET)
        return rk4_3field(p, f, 1);
    if (!rng) return -1;
    derive_consts(p, &k);
    double *star = (double *)malloc(sizeof(double) * n);
    if (!star) return -1;
    dornic_sweep(p, &k, rng, f, star); /* DRho == f[0..n), Rho_dt[0] <- star */
    switch (p->model) {
    case TMO_MODEL_DP: /* P/:1064-1068 */
        for (int64_t i = 0; i < n; i++) f[i] = star[i] / (1 + p->b * star[i] * p->dt);
        break;
    case TMO_MODEL_DP_CUBIC: { /* RK4_Integrate P/:1775-1790 */
        const double dt = p->dt, dt6 = dt / 6.0, dt2 = dt / 2.0, b = p->b, c = p->c;
        for (int64_t i = 0; i < n; i++) {
            const double r0 = star[i];
            double K1, K2, K3, K4, M;
            K1 = -b * r0 * r0 - c * r0 * r0 * r0;
            M = r0 + dt2 * K1;
            K2 = -b * M * M - c * M * M * M; M = r0 + dt2 * K2;
            K3 = -b * M * M - c * M * M * M; M = r0 + dt * K3;
            K4 = -b * M * M - c * M * M * M;
            f[i] = r0 + dt6 * (K1 + 2.0 * K2 + 2.0 * K3 + K4);
        }
        break;
    }
    default: { /* Rietkerk: R/:4139 then clamp+commit R/:4150-4181 (F/:4460 for the FPE bounds) */
        memcpy(f, star, sizeof(double) * n);
        if (rk4_3field(p, f, 0)) { free(star); return -1; }
        for (int64_t i = 0; i < 3 * n; i++)
            if (f[i] < p->clamp_eps && f[i] >= p->clamp_lo) f[i] = 0.0;
        break;
    }
    }
    free(star);
    return 0;
}

/* One full time step of the SPB = 2 (vegetation + grazer) runner, R/:3752-4183:
 *   vegetation Dornic R/:3752-3809 -> spatial means of DRho, vegetation cover factor R/:3813-3824 ->
 *   calc_gamma_2Sp_NonRefugia R/:3844 -> grazer Dornic R/:3908-4114 (skipped when <G> == 0, R/:3910) ->
 *   RK4_Integrate_Stochastic_2Sp R/:4137 -> clamp + commit R/:4150-4181.
 * f: [4][L*L] = P, G, W, O (DRho == Rho_dt on entry, as after a commit); gamma: the grazer's gamma array (state: the
 * reference never resets it); central: the full distance-sorted disc list of computeNeighboringSitesCentral(int(r_max));
 * fr_graz = pR[1]/(dx r_max_effective) (R/:3409).  veff (may be NULL): the grazer's advection vectors of this step. */
int tmo_step_2sp(const tmo_params *pveg, const tmo_graz_params *pg, const tmo_cons_params *pc, const int32_t *central, int n_central,
                 double fr_graz, tmo_rng *rng, double *f, double *gamma, double *veff, int trig_libm)
{
    return tmo_step_2sp_means(pveg, pg, pc, central, n_central, fr_graz, rng, f, gamma, veff, trig_libm, NULL);
}

/* The same step with the two spatial means {<P>, <G>} of DRho supplied by the caller (means != NULL) instead of the
 * left-to-right sums: the product sums in a fixed tree (a few ulp away), and some of the grazer's formulas are
 * ill-conditioned in gamma (D - dx/2 (1-gamma)|v|(...) cancels, R/:4008), so the tests compare bit for bit given the
 * product's two scalars and check those scalars against tmo_mean separately. */
int tmo_step_2sp_means(const tmo_params *pveg, const tmo_graz_params *pg, const tmo_cons_params *pc, const int32_t *central, int n_central,
                       double fr_graz, tmo_rng *rng, double *f, double *gamma, double *veff, int trig_libm, const double *means)
{
    if (!pveg || !pg || !pc || pveg->L != pg->L || pveg->L != pc->L) return -1;
    const int L = pveg->L;
    const int64_t n = (int64_t)L * L;
    double *work = malloc(sizeof(double) * (size_t)n * 8);
    if (!work) return -1;
    double *D = work, *vtmp = work + 4 * n, *gout = work + 6 * n, *vout = work + 7 * n;   /* DRho copy, v_eff, rho*_G, rho*_P */
    memcpy(D, f, sizeof(double) * (size_t)n * 4);
    dornic_consts k;
    derive_consts(pveg, &k);
    dornic_sweep(pveg, &k, rng, D, vout);                                               /* vegetation, R/:3752-3809 */
    const double avg_veg = means ? means[0] : tmo_mean(D, n), avg_graz = means ? means[1] : tmo_mean(D + n, n);   /* R/:3813-3818 */
    const double num_veg = tmo_occupied(D, n);
    double nR_fac = 1 - num_veg / (L * L);                                              /* R/:3822 */
    if (nR_fac < 0.35) nR_fac = 0.35;
    int new_length = (int)(nR_fac * nR_fac * n_central);                                /* R/:2539 */
    if (new_length < 1) new_length = 1;
    const int eff = (int)(fr_graz * fr_graz * new_length);                              /* R/:2569 */
    tmo_gamma_perception(L, D, central, eff, avg_veg, avg_graz, gamma);                 /* R/:3844 */
    if (avg_graz == 0.0) memcpy(gout, D + n, sizeof(double) * (size_t)n);               /* R/:3910 */
    else if (rng->kind == TMO_SRC_SEQ) {                                                /* the reference's single generator */
        if (tmo_grazer_dornic(pg, rng, D + n, gamma, gout, veff ? veff : vtmp, trig_libm)) { free(work); return -1; }
    } else {
        /* per-cell streams: the two species of a cell must not share one.  REPLAY: the grazer's slices follow the
         * vegetation's L*L slices in u; PHILOX: the grazer's key is seed ^ 0x9E3779B97F4A7C15 (the product's
         * kGrazerSeedSalt, tm_consumer.cuh). */
        tmo_rng gr = *rng;
        gr.overflow = 0; gr.draws = 0;
        if (rng->kind == TMO_SRC_REPLAY) {
            const uint64_t skip = (uint64_t)n * rng->stride;
            gr.u = rng->u + skip; gr.n_u = rng->n_u > skip ? rng->n_u - skip : 0;
        } else gr.seed = rng->seed ^ 0x9E3779B97F4A7C15ull;
        if (tmo_grazer_dornic(pg, &gr, D + n, gamma, gout, veff ? veff : vtmp, trig_libm)) { free(work); return -1; }
        rng->overflow += gr.overflow; rng->draws += gr.draws;
    }
    memcpy(f, vout, sizeof(double) * (size_t)n);
    memcpy(f + n, gout, sizeof(double) * (size_t)n);
    if (tmo_rk4_2sp(pc, f)) { free(work); return -1; }                                  /* R/:4137 */
    for (int64_t i = 0; i < 4 * n; i++)
        if (f[i] < pveg->clamp_eps && f[i] >= pveg->clamp_lo) f[i] = 0.0;               /* R/:4176 */
    free(work);
    return 0;
}

/* One time step of the SPB = 3 runner (vegetation + grazer + predator), rietkerk_Dornic_2D_MultiSp compiled with -DSPB=3, R/:3752-4183:
 * vegetation Dornic -> means of DRho[0..2] and the vegetation cover (R/:3813-3824) -> calc_gamma_3Sp_NonRefugia (R/:3850) -> the
 * consumers' Dornic loop for s = 1, 2 (R/:3908-4114; a species with <rho> == 0 is skipped, R/:3910) -> RK4_Integrate_Stochastic_MultiSp
 * (R/:4135) -> clamp (R/:4176).  Per-cell streams: the grazer's and the predator's keys are the seed xor 0x9E3779B97F4A7C15 and
 * 0xD1B54A32D192ED03 (PHILOX); their replay slices follow the vegetation's (REPLAY).  SEQ: the reference's single generator. */
int tmo_step_3sp(const tmo_params *pveg, const tmo_graz_params *pg1, const tmo_graz_params *pg2, const tmo_cons3_params *pc,
                 const int32_t *central, int n_central, double fr1, double fr2, tmo_rng *rng, double *f, double *gamma, double *veff,
                 int trig_libm, const double *means)
{
    if (!pveg || !pg1 || !pg2 || !pc || pveg->L != pg1->L || pveg->L != pg2->L || pveg->L != pc->L) return -1;
    const int L = pveg->L;
    const int64_t n = (int64_t)L * L;
    double *work = malloc(sizeof(double) * (size_t)n * 12);
    if (!work) return -1;
    double *D = work, *vtmp = work + 5 * n, *cout_ = work + 9 * n, *vout = work + 11 * n;   /* DRho copy, v_eff x2, rho* of G and Pr, rho*_P */
    memcpy(D, f, sizeof(double) * (size_t)n * 5);
    dornic_consts k;
    derive_consts(pveg, &k);
    dornic_sweep(pveg, &k, rng, D, vout);                                               /* vegetation, R/:3752-3809 */
    double avg[3];
    for (int s = 0; s < 3; s++) avg[s] = means ? means[s] : tmo_mean(D + s * n, n);      /* R/:3813-3818 */
    const double num_veg = tmo_occupied(D, n);
    double nR_fac = 1 - num_veg / (L * L);                                              /* R/:3822 */
    if (nR_fac < 0.35) nR_fac = 0.35;
    int new_length = (int)(nR_fac * nR_fac * n_central);                                /* R/:3162 */
    if (new_length < 1) new_length = 1;
    tmo_gamma_3sp(L, D, D + n, D + 2 * n, central, new_length, fr1, fr2, avg, gamma, gamma + n);   /* R/:3850 */
    const tmo_graz_params *pg[3] = {NULL, pg1, pg2};
    const uint64_t salt[3] = {0, 0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
    for (int s = 1; s <= 2; s++) {
        double *out = cout_ + (s - 1) * n, *ve = veff ? veff + 2 * (s - 1) * n : vtmp + 2 * (s - 1) * n;
        if (avg[s] == 0.0) { memcpy(out, D + s * n, sizeof(double) * (size_t)n); continue; }   /* R/:3910 */
        if (rng->kind == TMO_SRC_SEQ) {
            if (tmo_grazer_dornic(pg[s], rng, D + s * n, gamma + (s - 1) * n, out, ve, trig_libm)) { free(work); return -1; }
        } else {
            tmo_rng gr = *rng;
            gr.overflow = 0; gr.draws = 0;
            if (rng->kind == TMO_SRC_REPLAY) {
                const uint64_t skip = (uint64_t)s * n * rng->stride;
                gr.u = rng->u + skip; gr.n_u = rng->n_u > skip ? rng->n_u - skip : 0;
            } else gr.seed = rng->seed ^ salt[s];
            if (tmo_grazer_dornic(pg[s], &gr, D + s * n, gamma + (s - 1) * n, out, ve, trig_libm)) { free(work); return -1; }
            rng->overflow += gr.overflow; rng->draws += gr.draws;
        }
    }
    memcpy(f, vout, sizeof(double) * (size_t)n);
    memcpy(f + n, cout_, sizeof(double) * (size_t)n * 2);
    if (tmo_rk4_3sp(pc, f)) { free(work); return -1; }                                  /* R/:4135 */
    for (int64_t i = 0; i < 5 * n; i++)
        if (f[i] < pveg->clamp_eps && f[i] >= pveg->clamp_lo) f[i] = 0.0;               /* R/:4176 */
    free(work);
    return 0;
}

/* One time step of the Rietkerk_FPE runner compiled with -DSPB=3, rietkerk_DorFPE_2D_MultiSp F/:4094-4462 (see tm_oracle.h). */
int tmo_step_fpe_3sp(const tmo_params *pveg, const tmo_cons3_params *pc, const double *M, const double *v, const tmo_fke_terms *ft,
                     const double *const *gD, const double *const *gV, const int32_t *central, int n_central, double fr1, double fr2,
                     tmo_rng *rng, double *f, double *gamma, double *veff, int trig_libm, const double *means)
{
    if (!pveg || !pc || !ft || pveg->L != pc->L || pveg->L != ft[0].L || pveg->L != ft[1].L) return -1;
    const int L = pveg->L;
    const int64_t n = (int64_t)L * L;
    double *work = malloc(sizeof(double) * (size_t)n * 8);
    if (!work) return -1;
    double *D = work, *vout = work + 5 * n, *moved = work + 6 * n;                       /* DRho copy, rho*_P, moved consumers [2][n] */
    memcpy(D, f, sizeof(double) * (size_t)n * 5);
    double avg[3];
    for (int s = 0; s < 3; s++) avg[s] = means ? means[s] : tmo_mean(D + s * n, n);      /* F/:4101-4106 */
    const double num_veg = tmo_occupied(D, n);
    double nR_fac = 1 - num_veg / (L * L);                                               /* F/:4110 */
    if (nR_fac < 0.35) nR_fac = 0.35;
    int new_length = (int)(nR_fac * nR_fac * n_central);                                 /* F/:3380 */
    if (new_length < 1) new_length = 1;                                                  /* F/:3381 */
    /* --- calc_gamma_vel_3Sp_NonRefugia, F/:3371-3483 --- */
    {
        const double eps = 1.0e-12;
        double rho_inverse[3];
        for (int s = 0; s < 3; s++) rho_inverse[s] = avg[s] >= eps ? 1 / avg[s] : 1 / eps;
        const double fr[3] = {0.0, fr1, fr2}, vel[3] = {0.0, v[0], v[1]};
        int order[3] = {0, 1, 2};                                                        /* rfrac sorted by decreasing radius, F/:3745 */
        for (int a_ = 0; a_ < 3; a_++)
            for (int b_ = a_ + 1; b_ < 3; b_++)
                if (fr[order[b_]] > fr[order[a_]]) { const int t_ = order[a_]; order[a_] = order[b_]; order[b_] = t_; }
        double pi = 0.0, p16 = 1.0;                                                      /* the reference's PI, header :241-284 */
        for (int q = 0; q <= 14; q++) {
            pi = (1.0 / p16 * (4.0 / (8 * q + 1.0) - 2.0 / (8 * q + 4.0) - 1.0 / (8 * q + 5.0) - 1.0 / (8 * q + 6.0))) + pi;
            p16 *= 16.0;
        }
        tmo_rng dir = *rng;                                                              /* the direction uniforms' source */
        if (rng->kind == TMO_SRC_PHILOX) dir.seed = rng->seed ^ 0xA0761D6478BD642Full;
        else if (rng->kind == TMO_SRC_REPLAY) {
            const uint64_t skip = (uint64_t)n * rng->stride;
            dir.u = rng->u + skip; dir.n_u = rng->n_u > skip ? rng->n_u - skip : 0;
        }
        tmo_rng *dsrc = rng->kind == TMO_SRC_SEQ ? rng : &dir;
        if (dsrc == &dir) { dir.overflow = 0; dir.draws = 0; }
        cstream cs;
        cs_begin_cell(&cs, dsrc, 0);
        const double *fld[3] = {D, D + n, D + 2 * n};
        for (int i = 0; i < L * L; i++) {
            const int c_i = i / L, c_j = i % L;
            if (dsrc->kind != TMO_SRC_SEQ) cs_begin_cell(&cs, dsrc, (uint64_t)i);
            for (int o = 0; o < 3; o++) {
                const int s = order[o];
                int eff = (int)(fr[s] * fr[s] * new_length);                             /* F/:3413 */
                const double ran = next_u(&cs);                                          /* F/:3421: every species, the vegetation too */
                if (s >= 1) {
                    double vx, vy;
                    if (trig_libm) { vx = vel[s] * cos(2 * pi * ran); vy = vel[s] * sin(2 * pi * ran); }   /* F/:3424 */
                    else { double sn, csn; tmo_sincos_turn128(128.0 * ran, &sn, &csn); vx = vel[s] * csn; vy = vel[s] * sn; }
                    veff[2 * ((s - 1) * n + i)] = vx; veff[2 * ((s - 1) * n + i) + 1] = vy;
                }
                if (fr[s] == 0.0) continue;                                              /* F/:3426 */
                if (eff < 1 && fr[s] > 0) eff = 1;                                       /* F/:3428 */
                double *g = gamma + (s - 1) * n;
                if (avg[s] < eps) { g[i] = 1.0; continue; }                              /* F/:3430 */
                else if (avg[s - 1] < eps) { g[i] = 0.0; continue; }                     /* F/:3432 */
                const double *prey = fld[s - 1];
                if (s == 1) {
                    if (avg[2] < eps) {
                        for (int k = 0; k < eff; k++) {
                            const int nx = (c_i + central[2 * k] + L) % L, ny = (c_j + central[2 * k + 1] + L) % L;
                            g[i] += prey[nx * L + ny];                                   /* F/:3440: accumulated onto the old value */
                        }
                    } else {
                        for (int k = 0; k < eff; k++) {
                            const int nx = (c_i + central[2 * k] + L) % L, ny = (c_j + central[2 * k + 1] + L) % L;
                            g[i] += prey[nx * L + ny] * (1 - fld[2][nx * L + ny] * rho_inverse[2]);   /* F/:3447 */
                        }
                    }
                    g[i] = (g[i] * rho_inverse[s - 1]) / (eff);                          /* F/:3450 */
                } else {
                    for (int k = 0; k < eff; k++) {
                        const int nx = (c_i + central[2 * k] + L) % L, ny = (c_j + central[2 * k + 1] + L) % L;
                        g[i] += prey[nx * L + ny];
                    }
                    g[i] = (g[i] / (eff)) * rho_inverse[s - 1];                          /* F/:3456 */
                }
                if (g[i] > 1) g[i] = 1.0; else if (g[i] < 0) g[i] = 0.0;                 /* F/:3459-3462 */
            }
        }
        if (dsrc == &dir) { rng->overflow += dir.overflow; rng->draws += dir.draws; }
    }
    dornic_consts k;
    derive_consts(pveg, &k);
    dornic_sweep(pveg, &k, rng, D, vout);                                                /* vegetation only, F/:4154-4211 */
    memcpy(f, vout, sizeof(double) * (size_t)n);
    if (tmo_rk4_fpe_3sp(pc, M[0], M[1], f)) { free(work); return -1; }                   /* F/:4238 */
    for (int s = 1; s <= 2; s++) {                                                       /* advdiff_FKE_MultiSp, F/:4308 */
        if (tmo_fke_move(&ft[s - 1], f + s * n, gamma + (s - 1) * n, veff + 2 * (s - 1) * n, gD[s - 1], gV[s - 1], moved + (s - 1) * n)) { free(work); return -1; }
    }
    memcpy(f + n, moved, sizeof(double) * (size_t)n * 2);
    for (int64_t i = 0; i < 5 * n; i++)
        if (f[i] < ft[0].epsilon && f[i] >= 0.0) f[i] = 0.0;                             /* F/:4458 */
    free(work);
    return 0;
}

/* One time step of the Rietkerk_FPE runner compiled with -DSPB=2, rietkerk_DorFPE_2D_MultiSp F/:4094-4462 (see tm_oracle.h). */
int tmo_step_fpe_2sp(const tmo_params *pveg, const tmo_cons_params *pc, double M1, const tmo_fke_terms *ft, const double *gD,
                     const double *gV, const int32_t *central, int n_central, double fr_graz, tmo_rng *rng, double *f, double *gamma,
                     const double *veff, const double *means)
{
    if (!pveg || !pc || !ft || pveg->L != pc->L || pveg->L != ft->L) return -1;
    const int L = pveg->L;
    const int64_t n = (int64_t)L * L;
    double *work = malloc(sizeof(double) * (size_t)n * 6);
    if (!work) return -1;
    double *D = work, *vout = work + 4 * n, *moved = work + 5 * n;                       /* DRho copy, rho*_P, moved grazers */
    memcpy(D, f, sizeof(double) * (size_t)n * 4);
    const double avg_veg = means ? means[0] : tmo_mean(D, n), avg_graz = means ? means[1] : tmo_mean(D + n, n);   /* F/:4101-4106 */
    const double num_veg = tmo_occupied(D, n);
    double nR_fac = 1 - num_veg / (L * L);                                               /* F/:4110 */
    if (nR_fac < 0.35) nR_fac = 0.35;
    const int new_length = (int)(nR_fac * nR_fac * n_central);                           /* F/:2543 (no lower bound here) */
    const int eff = (int)(fr_graz * fr_graz * new_length);                               /* F/:2572 */
    tmo_gamma_perception_fpe(L, D, central, eff, avg_veg, avg_graz, gamma);              /* F/:4117 */
    dornic_consts k;
    derive_consts(pveg, &k);
    dornic_sweep(pveg, &k, rng, D, vout);                                                /* vegetation only, F/:4154-4211 */
    memcpy(f, vout, sizeof(double) * (size_t)n);
    if (tmo_rk4_fpe_2sp(pc, M1, f)) { free(work); return -1; }                           /* F/:4240 */
    if (tmo_fke_move(ft, f + n, gamma, veff, gD, gV, moved)) { free(work); return -1; }  /* F/:4303-4308 */
    memcpy(f + n, moved, sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < 4 * n; i++)
        if (f[i] < ft->epsilon && f[i] >= 0.0) f[i] = 0.0;                               /* F/:4458 */
    free(work);
    return 0;
}

/* ================================================================================== */
/* 6b. The noise-off DP path: percNoDiff_2D P/:549-662                                  */
/* ================================================================================== */
/* fP_2D P/:516-529 with determine_neighbours_R2 P/:394-414 ([0][j]: rows p-2..p+2 of column q; [1][j]: columns q-2..q+2
 * of row p; periodic) */
static void fP_2D(double *f, const double *R, int g, double a, double b, double D, double dx)
{
    for (int i = 0; i < g * g; i++) {
        const int p = i / g, q = i % g;
        int v[5], h[5];
        for (int j = 0; j < 5; j++) {
            const int k = j - 2;
            v[j] = ((p + k) % g + g) % g * g + q;
            h[j] = p * g + ((q + k) % g + g) % g;
        }
        f[i] = a * R[i] - b * R[i] * R[i] - D * (1 / (12.0 * dx * dx)) * (1 * (R[v[0]] + R[v[4]] + R[h[0]] + R[h[4]])
               - 16 * (R[v[1]] + R[v[3]] + R[h[1]] + R[h[3]]) + 60 * R[v[2]]);
    }
}

/* one RK4 step of percNoDiff_2D's time loop (P/:591-621) in place on rho[L*L] */
int tmo_dp_noise_off_step(int L, double a, double b, double D, double dt, double dx, double *rho)
{
    if (L < 1 || !rho) return -1;
    const int n = L * L;
    double *K1 = malloc(sizeof(double) * (size_t)n * 5), *K2 = K1 + n, *K3 = K2 + n, *K4 = K3 + n, *M = K4 + n;
    if (!K1) return -1;
    fP_2D(K1, rho, L, a, b, D, dx);
    for (int i = 0; i < n; i++) M[i] = rho[i] + (dt / 2.0) * K1[i];
    fP_2D(K2, M, L, a, b, D, dx);
    for (int i = 0; i < n; i++) M[i] = rho[i] + (dt / 2.0) * K2[i];
    fP_2D(K3, M, L, a, b, D, dx);
    for (int i = 0; i < n; i++) M[i] = rho[i] + (dt) * K3[i];
    fP_2D(K4, M, L, a, b, D, dx);
    for (int i = 0; i < n; i++) rho[i] += (dt / 6.0) * (K1[i] + 2 * K2[i] + 2 * K3[i] + K4[i]);
    free(K1);
    return 0;
}

/* ================================================================================== */
/* 7. Reductions and host-side bookkeeping                                            */
/* ================================================================================== */

double tmo_mean(const double *x, int64_t n) /* R/:1211-1219 = P/:123-131 */
{
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) sum += x[i];
    return sum / (double)n;
}

double tmo_occupied(const double *x, int64_t n) /* R/:1245-1254 = P/:145-154 */
{
    double sum = 0.0;
    for (int64_t i = 0; i < n; ++i) if (x[i] > 0) sum += 1;
    return sum;
}

int tmo_logarithm10_time_bins(double t_max, double dt, double *out, int cap) /* P/:336-378 */
{
    int n = 0;
    if (t_max <= 1) return 0;
    double tend_one = (double)(floor(log10(t_max) * 10.)) / 10.0;
    double power = 0.0;
    while (power <= tend_one) {
        double t_raw = pow(10, power);
        double t_fin = (int)(t_raw / dt) * dt;
        if (t_fin <= 1.0) t_fin = t_fin + dt;
        if (n > 0 && t_fin != out[n - 1]) { if (n < cap) out[n++] = t_fin; }
        else if (n == 0) { if (n < cap) out[n++] = t_fin; }
        power += 0.04;
    }
    for (int i = n - 1; i > 0; i--) {
        if (out[i] < out[i - 1] + dt / 2.0) {
            memmove(out + i, out + i + 1, sizeof(double) * (size_t)(n - i - 1));
            n--;
        }
    }
    return n;
}

void tmo_var_mean_incremental_surv_runs(double *acc, const double *rep, int n, int j) /* P/:173-208 */
{
    (void)j;
    for (int i = 0; i < n; i++) {
        double *A = acc + 5 * i; const double *X = rep + 2 * i;
        int r = (int)(A[3] + 1);
        if (X[0] == 0.0) break;
        double mean_prev = A[1];
        if (A[3] == 0.0) {
            A[1] = X[1]; A[2] = 0.0; A[4] = X[0];
        } else {
            A[1] = (X[1] + (r - 1) * A[1]) / r;
            A[4] = (X[0] + (r - 1) * A[4]) / r;
            A[2] = (A[1] - mean_prev) * (A[1] - mean_prev)
                 + (1 / (r - 1)) * ((r - 2) * A[2] + (X[1] - A[1]) * (X[1] - A[1])); /* int 1/(r-1) */
        }
        A[3] = r;
    }
}

/* spreading-experiment observables: meansq_spread_of_vector P/:156-169 (N = #{x>0}, sum over them of
 * (p-c_x)^2 + (q-c_y)^2 with p = i/g, q = i%g) and theborderwall P/:380-390. */
void tmo_spread_stats(const double *x, int g, int c_x, int c_y, double *sq_dist, int64_t *n_active, int *border)
{
    double sq = 0.0; int64_t sum = 0;
    for (int i = 0; i < g * g; i++)
        if (x[i] > 0) {
            int p = i / g, q = i % g;
            sum += 1;
            sq += (p - c_x) * (p - c_x) + (q - c_y) * (q - c_y);
        }
    *sq_dist = sq; *n_active = sum;
    int g2 = g * (g - 1), hit = 0;
    for (int i = 0; i < g && !hit; i++)
        if (x[i] > 0 || x[g * i] > 0 || x[g2 + i] > 0 || x[g * i + g - 1] > 0) hit = 1;
    *border = hit;
}

/* spreading_incremental_surv_runs P/:230-252: acc [n][4] = {t, <N>, <R2>, n_surv}; rep [n][3] = {N, R2, 1} */
void tmo_spreading_incremental_surv_runs(double *acc, const double *rep, int n)
{
    for (int i = 0; i < n; i++) {
        int r = (int)(acc[4 * i + 3] + 1);
        if (rep[3 * i] == 0.0) break;
        acc[4 * i + 1] = (rep[3 * i] + (r - 1) * acc[4 * i + 1]) / r;
        acc[4 * i + 2] = (rep[3 * i + 1] + (r - 1) * acc[4 * i + 2]) / r;
        acc[4 * i + 3] = r;
    }
}
