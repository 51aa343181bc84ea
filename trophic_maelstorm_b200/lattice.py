"""ctypes view of include/tm_lattice.h (no numerics live here).

Mirrors the C ABI one-to-one: ``Lattice`` owns a ``tm_ctx*``; arrays are numpy float64 in the
ABI's SoA order ``[n_trials][n_fields][nrows][L]``.
"""
import ctypes as C
import os

import numpy as np

MODEL_DP, MODEL_DP_CUBIC, MODEL_RIETKERK, MODEL_RIETKERK_DET = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    """The in-tree library (nothing else is ever loaded: no environment override, no fallback)."""
    return os.path.join(_HERE, "libtm_lattice.so")


class TmError(RuntimeError):
    pass


class Params(C.Structure):
    """struct tm_params (include/tm_lattice.h)."""
    _fields_ = [("model", C.c_int32), ("L", C.c_int32)] + [
        (n, C.c_double) for n in (
            "dt dx a b c D sigma rc gmax alpha rW W0 d K1 K2 D0 D1 D2 sigma0 clamp_eps clamp_lo").split()]


class Stats(C.Structure):
    """struct tm_stats."""
    _fields_ = [("mean", C.c_double * 3), ("active", C.c_uint64 * 3),
                ("nonfinite", C.c_uint64), ("negative", C.c_uint64)]


class Spread(C.Structure):
    """struct tm_spread."""
    _fields_ = [("n_active", C.c_uint64), ("sq_dist", C.c_double), ("border", C.c_int32), ("pad_", C.c_int32)]


def rietkerk_params(L, a=0.05, dt=0.1, model=MODEL_RIETKERK, fpe_clamp=False):
    """The SPB=1 constants of Rietkerk_FD/order_1stoc_test_rietkerk.cpp:14,28-37,48."""
    p = Params()
    p.model, p.L, p.dt, p.dx, p.a = model, L, dt, 0.1, a
    p.rc, p.gmax, p.alpha, p.rW, p.W0, p.d = 1e4, 0.05e-3 / 24, 0.2 / 24, 0.2 / 24, 0.2, 0.25 / 24
    p.K1, p.K2 = 5.0, 5000.0
    p.D0, p.D1, p.D2 = 0.00025 / 24.0, 0.00025 / 24.0, 0.025 / 24.0
    p.sigma0 = (p.D0 / (p.dx * p.dx)) ** 0.5
    p.clamp_eps, p.clamp_lo = (1e-10, 0.0) if fpe_clamp else (1e-12, -1e-80)
    return p


def dp_params(L, a, b, D, sigma, dt, dx, c=None):
    """DP (c is None) / cubic DP constants as the Percolation_FD drivers pass them."""
    p = Params()
    p.model = MODEL_DP if c is None else MODEL_DP_CUBIC
    p.L, p.dt, p.dx, p.a, p.b, p.c, p.D, p.sigma = L, dt, dx, a, b, (0.0 if c is None else c), D, sigma
    return p


_lib = None


def lib():
    """Load libtm_lattice.so; raises TmError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise TmError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a). There is no CPU fallback." % path)
    L = C.CDLL(path)
    vp, dp, i, u64 = C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_uint64
    sig = {
        "tm_abi_version": (i, []),
        "tm_device_count": (i, []),
        "tm_device_mem": (i, [i, C.POINTER(u64), C.POINTER(u64)]),
        "tm_ctx_bytes": (i, [C.POINTER(Params), i, C.POINTER(u64)]),
        "tm_last_error": (C.c_char_p, [vp]),
        "tm_create": (i, [C.POINTER(Params), i, i, C.POINTER(vp)]),
        "tm_create_slab": (i, [C.POINTER(Params), i, i, i, i, C.POINTER(vp)]),
        "tm_destroy": (i, [vp]),
        "tm_set_trial_a": (i, [vp, i, C.c_double]),
        "tm_set_strict": (i, [vp, i]),
        "tm_set_fused": (i, [vp, i]),
        "tm_upload": (i, [vp, dp]),
        "tm_download": (i, [vp, dp]),
        "tm_n_fields": (i, [vp]),
        "tm_seed": (i, [vp, u64, C.c_uint32]),
        "tm_set_step_counter": (i, [vp, u64]),
        "tm_step_counter": (u64, [vp]),
        "tm_step": (i, [vp, i]),
        "tm_step_replay": (i, [vp, dp, u64]),
        "tm_dornic_only": (i, [vp, dp]),
        "tm_halo_rows": (i, []),
        "tm_halo_bytes": (i, [vp, C.POINTER(u64)]),
        "tm_halo_pack": (i, [vp, i, C.POINTER(vp)]),
        "tm_halo_unpack": (i, [vp, i, vp]),
        "tm_cuda_stream": (vp, [vp]),
        "tm_slab_band_rows": (i, []),
        "tm_state_buffers": (i, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        "tm_slab_comm_stream": (vp, [vp]),
        "tm_slab_reserve_sms": (i, [vp, i]),
        "tm_slab_halo_ptrs": (i, [vp, i, i, i, i, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]),
        "tm_slab_step_begin": (i, [vp]),
        "tm_slab_step_end": (i, [vp]),
        "tm_slab_join": (i, [vp]),
        "tm_slab_pull_halos": (i, [vp, vp, vp, i]),
        "tm_dp_noise_off_host": (i, [i, i] + [C.c_double] * 5 + [dp, i, C.POINTER(C.c_float)]),
        "tm_measure": (i, [vp, C.POINTER(Stats)]),
        "tm_measure_spread": (i, [vp, i, i, C.POINTER(Spread)]),
        "tm_sampler_overflow": (i, [vp, C.POINTER(u64)]),
        "tm_sync": (i, [vp]),
        "tm_last_step_ms": (i, [vp, C.POINTER(C.c_float), C.POINTER(i)]),
        "tm_set_kernel_timing": (i, [vp, i]),
        "tm_kernel_ms": (i, [vp, C.POINTER(C.c_double), C.POINTER(i)]),
        "tm_run_host": (i, [vp, dp, i, i, C.POINTER(Stats)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


ABI_SYMBOLS = [
    "tm_abi_version", "tm_device_count", "tm_device_mem", "tm_ctx_bytes", "tm_last_error", "tm_create", "tm_create_slab", "tm_destroy",
    "tm_set_trial_a", "tm_set_strict", "tm_set_fused", "tm_upload", "tm_download", "tm_n_fields", "tm_seed",
    "tm_set_step_counter", "tm_step_counter", "tm_step", "tm_step_replay", "tm_dornic_only",
    "tm_halo_rows", "tm_halo_bytes", "tm_halo_pack", "tm_halo_unpack", "tm_cuda_stream",
    "tm_slab_band_rows", "tm_state_buffers", "tm_slab_comm_stream", "tm_slab_reserve_sms", "tm_slab_halo_ptrs", "tm_slab_step_begin", "tm_slab_step_end",
    "tm_slab_join", "tm_slab_pull_halos", "tm_measure", "tm_measure_spread",
    "tm_sampler_overflow", "tm_sync", "tm_last_step_ms", "tm_set_kernel_timing", "tm_kernel_ms", "tm_run_host",
    "tm_dp_noise_off_host",
]


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def dp_noise_off(L, a, b, D, dt, dx, rho, nsteps, device=0):
    """tm_dp_noise_off_host: nsteps RK4 steps of the noise-off DP path (percNoDiff_2D) -> (rho', device ms)."""
    r = np.ascontiguousarray(rho, dtype=np.float64).copy()
    assert r.size == L * L
    ms = C.c_float()
    rc = lib().tm_dp_noise_off_host(device, L, a, b, D, dt, dx, _dptr(r), nsteps, C.byref(ms))
    if rc != 0:
        raise TmError("tm_dp_noise_off_host failed (%d): %s" % (rc, lib().tm_last_error(None).decode()))
    return r, ms.value


class Lattice:
    """One tm_ctx: n_trials lattices of one model on one device."""

    def __init__(self, params, n_trials=1, device=0, row0=0, nrows=None):
        self._l = lib()
        self.params = params
        self.n_trials = n_trials
        self.L = params.L
        self.nrows = params.L if nrows is None else nrows
        self._ctx = C.c_void_p()
        rc = self._l.tm_create_slab(C.byref(params), n_trials, device, row0, self.nrows, C.byref(self._ctx))
        if rc != 0:
            self._ctx = C.c_void_p()
            raise TmError("tm_create failed (%d): %s" % (rc, self._l.tm_last_error(None).decode()))
        self.n_fields = self._l.tm_n_fields(self._ctx)

    def _ck(self, rc):
        if rc != 0:
            raise TmError("tm_lattice call failed (%d): %s" % (rc, self._l.tm_last_error(self._ctx).decode()))

    @property
    def shape(self):
        return (self.n_trials, self.n_fields, self.nrows, self.L)

    def close(self):
        if self._ctx:
            self._l.tm_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_trial_a(self, trial, a):
        self._ck(self._l.tm_set_trial_a(self._ctx, trial, a))

    def set_strict(self, strict=True):
        self._ck(self._l.tm_set_strict(self._ctx, 1 if strict else 0))

    def set_fused(self, fused=True):
        self._ck(self._l.tm_set_fused(self._ctx, 1 if fused else 0))

    def upload(self, fields):
        a = np.ascontiguousarray(fields, dtype=np.float64)
        assert a.size == int(np.prod(self.shape)), (a.shape, self.shape)
        self._ck(self._l.tm_upload(self._ctx, _dptr(a)))
        self._ck(self._l.tm_sync(self._ctx))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        self._ck(self._l.tm_download(self._ctx, _dptr(out)))
        return out

    def seed(self, seed, trial0=0):
        self._ck(self._l.tm_seed(self._ctx, seed, trial0))

    def set_step_counter(self, step):
        self._ck(self._l.tm_set_step_counter(self._ctx, step))

    def step(self, nsteps=1):
        self._ck(self._l.tm_step(self._ctx, nsteps))

    def step_replay(self, uniforms, stride):
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        assert u.size == self.n_trials * self.L * self.L * stride
        self._ck(self._l.tm_step_replay(self._ctx, _dptr(u), stride))

    def dornic_only(self):
        out = np.empty((self.n_trials, self.nrows, self.L), dtype=np.float64)
        self._ck(self._l.tm_dornic_only(self._ctx, _dptr(out)))
        return out

    def measure(self):
        st = (Stats * self.n_trials)()
        self._ck(self._l.tm_measure(self._ctx, st))
        return st

    def measure_spread(self, c_x, c_y):
        st = (Spread * self.n_trials)()
        self._ck(self._l.tm_measure_spread(self._ctx, int(c_x), int(c_y), st))
        return st

    def sampler_overflow(self):
        n = C.c_uint64()
        self._ck(self._l.tm_sampler_overflow(self._ctx, C.byref(n)))
        return n.value

    def sync(self):
        self._ck(self._l.tm_sync(self._ctx))

    def last_step_ms(self):
        ms, n = C.c_float(), C.c_int()
        self._ck(self._l.tm_last_step_ms(self._ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_kernel_timing(self, on=True):
        self._ck(self._l.tm_set_kernel_timing(self._ctx, 1 if on else 0))

    def kernel_ms(self):
        """(ms of the Dornic sweeps, ms of the RK4 kernels, steps covered) of the last step() call."""
        ms, n = (C.c_double * 2)(), C.c_int()
        self._ck(self._l.tm_kernel_ms(self._ctx, ms, C.byref(n)))
        return ms[0], ms[1], n.value

    def run_host(self, fields, nsteps, download=True):
        st = (Stats * self.n_trials)()
        self._ck(self._l.tm_run_host(self._ctx, _dptr(fields), nsteps, 1 if download else 0, st))
        return st

    def halo_bytes(self):
        n = C.c_uint64()
        self._ck(self._l.tm_halo_bytes(self._ctx, C.byref(n)))
        return n.value

    def halo_pack(self, which):
        p = C.c_void_p()
        self._ck(self._l.tm_halo_pack(self._ctx, which, C.byref(p)))
        return p.value

    def halo_unpack(self, which, dev_ptr):
        self._ck(self._l.tm_halo_unpack(self._ctx, which, C.c_void_p(dev_ptr)))

    def cuda_stream(self):
        return self._l.tm_cuda_stream(self._ctx)
