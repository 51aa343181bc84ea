"""The sampler spec: Philox known answers, the bit-reproducible elementary functions,
and the Poisson / Gamma / normal laws of the oracle's restatement of libstdc++ (LS)."""
import ctypes as C
import math

import numpy as np
import pytest
from scipy import stats

import oracle_lib as O


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    L = O.oracle()
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        out = (C.c_uint32 * 4)()
        L.tmo_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == want


def test_u52_uniform():
    L = O.oracle()
    assert L.tmo_u52_from_bits(0) == 0.0
    assert L.tmo_u52_from_bits(0xFFFFFFFFFFFFFFFF) == 1.0 - 2.0 ** -52
    assert L.tmo_u52_from_bits(1 << 63) == 0.5


def _ulps(a, b):
    return abs(a - b) / np.spacing(abs(b))


def test_log_exp_lgamma_accuracy():
    """<= 1 ulp vs libm for log/exp, <= 4 ulp for lgamma at integers (libm's own error included)."""
    L = O.oracle()
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.random(20000), 10 ** rng.uniform(-32, 3, 20000), rng.uniform(0.5, 2, 20000)])
    assert max(_ulps(L.tmo_log(float(x)), math.log(x)) for x in xs) <= 1.0
    xs = np.concatenate([rng.uniform(-700, 5, 30000), rng.uniform(-1, 1, 20000)])
    assert max(_ulps(L.tmo_exp(float(x)), math.exp(x)) for x in xs) <= 1.0
    assert L.tmo_exp(-800.0) == 0.0 and L.tmo_exp(-745.0) == 5e-324
    assert max(_ulps(L.tmo_lgamma_int(float(k)), math.lgamma(k)) for k in range(3, 3000)) <= 4.0
    assert L.tmo_lgamma_int(1.0) == 0.0 and L.tmo_lgamma_int(2.0) == 0.0


def _draws(fn, n, *args, seed=0):
    rng = np.random.default_rng(seed)
    u = rng.random(64 * n)
    pos = C.c_uint64(0)
    out = []
    for _ in range(n):
        out.append(fn(O.ptr(u), u.size, C.byref(pos), *args))
    assert pos.value < u.size
    return np.array(out)


@pytest.mark.parametrize("mean", [1e-3, 1.0, 11.9, 12.0, 37.5, 100.0])
def test_poisson_law(mean):
    """LS:1310-1400: product method below 12, Devroye rejection from 12 (chi-square vs the pmf)."""
    L = O.oracle()
    x = _draws(L.tmo_test_poisson, 40000, C.c_double(mean), seed=int(mean * 1000) + 1).astype(int)
    assert abs(x.mean() - mean) < 5 * math.sqrt(mean / x.size) + 1e-12
    lo, hi = int(stats.poisson.ppf(1e-4, mean)), int(stats.poisson.ppf(1 - 1e-4, mean))
    ks = np.arange(lo, hi + 1)
    exp = stats.poisson.pmf(ks, mean) * x.size
    obs = np.array([(x == k).sum() for k in ks], dtype=float)
    keep = exp >= 5
    if keep.sum() >= 2:
        o, e = obs[keep], exp[keep]
        e = e * o.sum() / e.sum()
        assert stats.chisquare(o, e).pvalue > 1e-4


@pytest.mark.parametrize("mean", [1e-3, 0.3, 1.0, 5.0, 11.9, 12.0, 30.0, 100.0, 1e4])
def test_poisson_law_philox(mean):
    """Philox streams draw Poisson by inversion (< 12) and PTRS transformed rejection (>= 12): same pmf (chi-square)."""
    L = O.oracle()
    n = 40000
    x = np.array([L.tmo_test_poisson_philox(1234, cell, 2, int(mean * 7) % 1000, C.c_double(mean)) for cell in range(n)])
    assert (x >= 0).all()
    assert abs(x.mean() - mean) < 5 * math.sqrt(mean / n) + 1e-12
    assert abs(x.var() - mean) < 0.05 * mean + 5e-3
    lo, hi = int(stats.poisson.ppf(1e-4, mean)), int(stats.poisson.ppf(1 - 1e-4, mean))
    ks = np.arange(lo, hi + 1)
    exp = stats.poisson.pmf(ks, mean) * n
    obs = np.array([(x == k).sum() for k in ks], dtype=float)
    keep = exp >= 5
    if keep.sum() >= 2:
        o, e = obs[keep], exp[keep]
        e = e * o.sum() / e.sum()
        assert stats.chisquare(o, e).pvalue > 1e-4


@pytest.mark.parametrize("shape", [1e-3, 0.5, 1.0, 2.5, 10.0, 100.0])
def test_gamma_law(shape):
    """LS:2352-2387: Marsaglia-Tsang, u^(1/alpha) boost below 1 (KS vs the cdf)."""
    L = O.oracle()
    x = _draws(L.tmo_test_gamma, 20000, C.c_double(shape), seed=int(shape * 1000) + 7)
    assert (x >= 0).all()
    if shape >= 0.5:
        assert stats.kstest(x, "gamma", args=(shape,)).pvalue > 1e-4
    else:  # nearly all mass at ~0: compare the mean and the upper tail instead
        assert abs(x.mean() - shape) < 6 * math.sqrt(shape / x.size)


def test_gamma_shape_zero_returns_zero():
    """alpha == 0 (a lone live cell that draws Poisson 0): LS returns u^(inf) * ... = 0 and burns draws."""
    L = O.oracle()
    u = np.random.default_rng(3).random(64)
    pos = C.c_uint64(0)
    assert L.tmo_test_gamma(O.ptr(u), u.size, C.byref(pos), C.c_double(0.0)) == 0.0
    assert pos.value >= 3


def test_normal_law_and_cache_order():
    """LS:1811-1848: polar method; first return is y*mult, the cached one x*mult."""
    L = O.oracle()
    rng = np.random.default_rng(11)
    u = rng.random(200000)
    pos = C.c_uint64(0)
    a, b = [], []
    while pos.value < u.size - 200:
        second = C.c_double()
        a.append(L.tmo_test_normal_pair(O.ptr(u), u.size, C.byref(pos), C.byref(second)))
        b.append(second.value)
    z = np.array(a + b)
    assert stats.kstest(z, "norm").pvalue > 1e-4
    # explicit check of the first accepted pair
    pos = C.c_uint64(0)
    second = C.c_double()
    first = L.tmo_test_normal_pair(O.ptr(u), u.size, C.byref(pos), C.byref(second))
    k = 0
    while True:
        x, y = 2 * u[k] - 1, 2 * u[k + 1] - 1
        k += 2
        r2 = x * x + y * y
        if 0 < r2 <= 1:
            break
    mult = math.sqrt(-2 * math.log(r2) / r2)
    assert pos.value == k
    assert abs(first - y * mult) <= 4e-16 * abs(first) and abs(second.value - x * mult) <= 4e-16 * abs(second.value)


def _sincos_exact(a):
    """sin, cos of 2 pi a/128 for a double a, in 50-digit decimal arithmetic (a is used exactly)."""
    from decimal import Decimal, getcontext
    getcontext().prec = 50
    pi = Decimal("3.14159265358979323846264338327950288419716939937510582097494")
    q = round(a / 32.0)                                   # quadrant; the remainder is exact
    x = 2 * pi * (Decimal(a) - 32 * q) / 128
    s, c, t = Decimal(0), Decimal(0), Decimal(1)
    for n in range(40):
        if n % 2 == 0:
            c += t if (n // 2) % 2 == 0 else -t
        else:
            s += t if (n // 2) % 2 == 0 else -t
        t = t * x / (n + 1)
    return [(s, c), (c, -s), (-s, -c), (-c, s)][q % 4]


def test_sincos_turn128_accuracy():
    """The Box-Muller angle: sin/cos of 2 pi a/128 on [0,128]: within 2^-52 of the exact value, exact on the axes,
    and the table the product carries equals the oracle's."""
    L = O.oracle()
    rng = np.random.default_rng(5)
    a = np.concatenate([rng.random(4000) * 128, np.arange(0, 128.01, 0.5), 128 * (1 - 2.0 ** -np.arange(1, 53))])
    sn, cs = C.c_double(), C.c_double()
    worst = 0.0
    for x in a:
        L.tmo_sincos_turn128(float(x), C.byref(sn), C.byref(cs))
        ws, wc = _sincos_exact(float(x))
        worst = max(worst, abs(sn.value - float(ws)), abs(cs.value - float(wc)))
        assert abs(sn.value ** 2 + cs.value ** 2 - 1) < 6e-16
    assert worst <= 2.0 ** -52
    for q, (ws, wc) in enumerate([(0.0, 1.0), (1.0, 0.0), (0.0, -1.0), (-1.0, 0.0), (0.0, 1.0)]):
        L.tmo_sincos_turn128(32.0 * q, C.byref(sn), C.byref(cs))
        assert (sn.value, cs.value) == (ws, wc)
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    strip = lambda p: [l.split("/*")[0].strip() for l in open(p) if l.lstrip().startswith(("0x", "-0x"))]
    assert strip(os.path.join(root, "oracle", "tm_sincos_table.inc")) == \
        strip(os.path.join(root, "trophic_maelstorm_b200", "csrc", "tm_sincos_table.inc"))


def test_box_muller_law_philox():
    """Philox streams draw N(0,1) by Box-Muller (no rejection loop): law, pair independence, and the
    explicit formula on the two uniforms of block 0."""
    L = O.oracle()
    a, b = [], []
    second = C.c_double()
    for cell in range(60000):
        a.append(L.tmo_test_normal_pair_philox(77, cell, 3, 9, C.byref(second)))
        b.append(second.value)
    a, b = np.array(a), np.array(b)
    assert stats.kstest(np.concatenate([a, b]), "norm").pvalue > 1e-4
    assert stats.kstest(a * a + b * b, "chi2", args=(2,)).pvalue > 1e-4
    assert abs(np.corrcoef(a, b)[0, 1]) < 0.02 and abs(np.corrcoef(a * a, b * b)[0, 1]) < 0.02
    assert np.isfinite(a).all() and np.isfinite(b).all()
    out = (C.c_uint32 * 4)()
    L.tmo_philox4x32_10((C.c_uint32 * 4)(5, 3, 9, 0), (C.c_uint32 * 2)(77, 0), out)
    u0 = L.tmo_u52_from_bits((out[1] << 32) | out[0])
    u1 = L.tmo_u52_from_bits((out[3] << 32) | out[2])
    x = 1.0 - u0 - (0.0 if (int((1.0 + u0) * 2 ** 52) & 1) else 2.0 ** -52)   # last mantissa bit of 1+u0 set
    r = math.sqrt(-2 * math.log(x))
    first = L.tmo_test_normal_pair_philox(77, 5, 3, 9, C.byref(second))
    assert abs(first - r * math.cos(2 * math.pi * u1)) <= 1e-14 and abs(second.value - r * math.sin(2 * math.pi * u1)) <= 1e-14


def test_incremental_stats_integer_division_quirk():
    """var_mean_incremental_surv_runs (P/:173-208): `1/(r-1)` is integer division, so the second
    variance term survives only for r == 2."""
    L = O.oracle()
    acc = np.zeros((1, 5)); acc[0, 1], acc[0, 3], acc[0, 4] = 1.0, 1, 10
    rep = np.array([[12.0, 3.0]])
    L.tmo_var_mean_incremental_surv_runs(O.ptr(acc), O.ptr(rep), 1, 1)     # r becomes 2
    assert acc[0, 1] == 2.0 and acc[0, 3] == 2 and acc[0, 4] == 11.0
    assert acc[0, 2] == (2.0 - 1.0) ** 2 + 1 * (0 * 0.0 + (3.0 - 2.0) ** 2)
    rep = np.array([[8.0, 5.0]])
    L.tmo_var_mean_incremental_surv_runs(O.ptr(acc), O.ptr(rep), 1, 2)     # r becomes 3: 1/(3-1) == 0
    assert acc[0, 1] == 3.0 and acc[0, 2] == (3.0 - 2.0) ** 2
