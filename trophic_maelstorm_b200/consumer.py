"""ctypes view of include/tm_consumer.h (no numerics live here): consumer-side operators of the SPB = 2 model."""
import ctypes as C

import numpy as np

from .lattice import Params, TmError, lib


class ConsParams(C.Structure):
    """struct tm_cons_params."""
    _fields_ = [("L", C.c_int32), ("pad_", C.c_int32)] + [(n, C.c_double) for n in "dt a c gmax alpha rW W0 K1 K2 Dw Do A01 H01 E1".split()]


class GrazParams(C.Structure):
    """struct tm_graz_params."""
    _fields_ = [("L", C.c_int32), ("pad_", C.c_int32)] + [(n, C.c_double) for n in "dt dx D v M sigma D0_floor".split()]


CONS_ABI_SYMBOLS = ["tm_cons_create", "tm_cons_destroy", "tm_cons_last_error", "tm_cons_rk4_host", "tm_cons_gamma_host",
                    "tm_cons_grazer_dornic_host", "tm_cons_last_ms", "tm_cons_step2_setup", "tm_cons_step2_upload", "tm_cons_step2_run",
                    "tm_cons_step2_replay", "tm_cons_step2_download", "tm_cons_step2_set_rk4_stages", "tm_cons_step2_set_fpe", "tm_cons_step2_set_strict_means"]

_bound = False


def _lib():
    global _bound
    L = lib()
    if not _bound:
        vp, dp, i = C.c_void_p, C.POINTER(C.c_double), C.c_int
        sig = {
            "tm_cons_create": (i, [i, i, C.POINTER(vp)]),
            "tm_cons_destroy": (i, [vp]),
            "tm_cons_last_error": (C.c_char_p, [vp]),
            "tm_cons_rk4_host": (i, [vp, C.POINTER(ConsParams), dp, i]),
            "tm_cons_gamma_host": (i, [vp, dp, C.POINTER(C.c_int32), i, C.c_double, C.c_double, dp]),
            "tm_cons_grazer_dornic_host": (i, [vp, C.POINTER(GrazParams), C.c_uint64, C.c_uint32, C.c_uint32, dp, C.c_uint32, dp, dp, dp, dp,
                                               C.POINTER(C.c_uint64)]),
            "tm_cons_last_ms": (i, [vp, C.POINTER(C.c_float), C.POINTER(i)]),
            "tm_cons_step2_setup": (i, [vp, C.POINTER(Params), C.POINTER(GrazParams), C.POINTER(ConsParams), C.POINTER(C.c_int32), i, C.c_double]),
            "tm_cons_step2_upload": (i, [vp, dp, dp]),
            "tm_cons_step2_run": (i, [vp, C.c_uint64, C.c_uint32, C.c_uint32, i]),
            "tm_cons_step2_replay": (i, [vp, dp, dp, C.c_uint32]),
            "tm_cons_step2_set_rk4_stages": (i, [vp, i]),
            "tm_cons_step2_set_strict_means": (i, [vp, i]),
            "tm_cons_step2_set_fpe": (i, [vp, vp, dp, dp, C.c_double]),
            "tm_cons_step2_download": (i, [vp, dp, dp, dp, dp, C.POINTER(C.c_uint64)]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _bound = True
    return L


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class ConsumerOps:
    """One tm_cons_ctx: perception field and four-field RK4 of the vegetation + grazer model on an L x L lattice."""

    def __init__(self, L, device=0):
        self._l = _lib()
        self.L = L
        self._ctx = C.c_void_p()
        rc = self._l.tm_cons_create(L, device, C.byref(self._ctx))
        if rc != 0:
            self._ctx = C.c_void_p()
            raise TmError("tm_cons_create failed (%d): %s" % (rc, self._l.tm_cons_last_error(None).decode()))

    def _ck(self, rc):
        if rc != 0:
            raise TmError("tm_cons call failed (%d): %s" % (rc, self._l.tm_cons_last_error(self._ctx).decode()))

    def close(self):
        if self._ctx:
            self._l.tm_cons_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rk4(self, params, fields, nsteps=1):
        f = np.ascontiguousarray(fields, dtype=np.float64).copy()
        assert f.size == 4 * self.L * self.L
        self._ck(self._l.tm_cons_rk4_host(self._ctx, C.byref(params), _dptr(f), nsteps))
        return f

    def gamma(self, prey, offsets, avg_prey, avg_self, gamma0):
        g = np.ascontiguousarray(gamma0, dtype=np.float64).copy()
        prey = np.ascontiguousarray(prey, dtype=np.float64)
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        self._ck(self._l.tm_cons_gamma_host(self._ctx, _dptr(prey), off.ctypes.data_as(C.POINTER(C.c_int32)), len(off), avg_prey, avg_self, _dptr(g)))
        return g

    def grazer_dornic(self, params, prev, gamma, seed=0, trial=0, step=0, replay=None, stride=0):
        """-> (rho*, gamma clamped, v_eff [L*L, 2], overflow count)."""
        n = self.L * self.L
        prev = np.ascontiguousarray(prev, dtype=np.float64)
        g = np.ascontiguousarray(gamma, dtype=np.float64).copy()
        out, veff, ovf = np.empty(n), np.empty((n, 2)), C.c_uint64()
        ru = None
        if replay is not None:
            replay = np.ascontiguousarray(replay, dtype=np.float64)
            assert replay.size == n * stride
            ru = _dptr(replay)
        self._ck(self._l.tm_cons_grazer_dornic_host(self._ctx, C.byref(params), seed, trial, step, ru, stride, _dptr(prev), _dptr(g),
                                                    _dptr(out), _dptr(veff), C.byref(ovf)))
        return out, g, veff, ovf.value

    # ---- the SPB = 2 time step resident on the device (tm_cons_step2_*) ----
    def step2_setup(self, veg, graz, cons, central, fr):
        cen = np.ascontiguousarray(central, dtype=np.int32)
        self._ck(self._l.tm_cons_step2_setup(self._ctx, C.byref(veg), C.byref(graz), C.byref(cons), cen.ctypes.data_as(C.POINTER(C.c_int32)),
                                             len(cen), fr))

    def step2_upload(self, fields, gamma=None):
        f = np.ascontiguousarray(fields, dtype=np.float64)
        assert f.size == 4 * self.L * self.L
        g = None if gamma is None else np.ascontiguousarray(gamma, dtype=np.float64)
        self._ck(self._l.tm_cons_step2_upload(self._ctx, _dptr(f), None if g is None else _dptr(g)))

    def step2_set_rk4_stages(self, on):
        self._ck(self._l.tm_cons_step2_set_rk4_stages(self._ctx, 1 if on else 0))

    def step2_set_strict_means(self, on):
        self._ck(self._l.tm_cons_step2_set_strict_means(self._ctx, 1 if on else 0))

    def step2_set_fpe(self, terms, gauss_D, gauss_V, M1):
        """Rietkerk_FPE flavour of the step (tm_cons_step2_set_fpe); terms: fke.FkeTerms, or None for the Rietkerk_FD flavour."""
        if terms is None:
            self._ck(self._l.tm_cons_step2_set_fpe(self._ctx, None, None, None, 0.0))
            return
        gD, gV = np.ascontiguousarray(gauss_D, dtype=np.float64), np.ascontiguousarray(gauss_V, dtype=np.float64)
        assert gD.size >= terms.n_gauss_D and gV.size >= terms.n_gauss_V
        self._ck(self._l.tm_cons_step2_set_fpe(self._ctx, C.cast(C.byref(terms), C.c_void_p), _dptr(gD), _dptr(gV), float(M1)))

    def step2_run(self, seed, trial=0, step0=0, nsteps=1):
        self._ck(self._l.tm_cons_step2_run(self._ctx, seed, trial, step0, nsteps))

    def step2_replay(self, u_veg, u_graz, stride):
        n = self.L * self.L
        uv, ug = np.ascontiguousarray(u_veg, dtype=np.float64), np.ascontiguousarray(u_graz, dtype=np.float64)
        assert uv.size == n * stride and ug.size == n * stride
        self._ck(self._l.tm_cons_step2_replay(self._ctx, _dptr(uv), _dptr(ug), stride))

    def step2_download(self):
        """-> (fields [4, L*L], gamma, v_eff [L*L, 2], means {<P>, <G>, #{P>0}, eff} of the last step's entry state, overflow)."""
        n = self.L * self.L
        f, g, v, m, ovf = np.empty((4, n)), np.empty(n), np.empty((n, 2)), np.empty(4), C.c_uint64()
        self._ck(self._l.tm_cons_step2_download(self._ctx, _dptr(f), _dptr(g), _dptr(v), _dptr(m), C.byref(ovf)))
        return f, g, v, m, ovf.value

    def last_ms(self):
        ms, n = C.c_float(), C.c_int()
        self._ck(self._l.tm_cons_last_ms(self._ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value


# ---- SPB = 3 (vegetation + grazer + predator): include/tm_consumer3.h ---------------------------------------------------------------
class Cons3Params(C.Structure):
    """struct tm_cons3_params."""
    _fields_ = [("L", C.c_int32), ("pad_", C.c_int32)] + [(n, C.c_double) for n in
                                                         "dt a c gmax alpha rW W0 K1 K2 Dw Do A01 H01 E1 A12 H12 E2".split()]


CONS3_ABI_SYMBOLS = ["tm_cons3_create", "tm_cons3_destroy", "tm_cons3_last_error", "tm_cons3_setup", "tm_cons3_set_fpe", "tm_cons3_upload", "tm_cons3_run",
                     "tm_cons3_replay", "tm_cons3_download", "tm_cons3_last_ms", "tm_cons3_set_strict_means"]
_bound3 = False


def _lib3():
    global _bound3
    L = lib()
    if not _bound3:
        vp, dp, i = C.c_void_p, C.POINTER(C.c_double), C.c_int
        sig = {
            "tm_cons3_create": (i, [i, i, C.POINTER(vp)]),
            "tm_cons3_destroy": (i, [vp]),
            "tm_cons3_last_error": (C.c_char_p, [vp]),
            "tm_cons3_setup": (i, [vp, C.POINTER(Params), C.POINTER(GrazParams), C.POINTER(GrazParams), C.POINTER(Cons3Params),
                                   C.POINTER(C.c_int32), i, C.c_double, C.c_double]),
            "tm_cons3_set_fpe": (i, [vp, vp, C.POINTER(dp), C.POINTER(dp), C.c_double, C.c_double]),
            "tm_cons3_upload": (i, [vp, dp, dp]),
            "tm_cons3_set_strict_means": (i, [vp, i]),
            "tm_cons3_run": (i, [vp, C.c_uint64, C.c_uint32, C.c_uint32, i]),
            "tm_cons3_replay": (i, [vp, dp, dp, dp, C.c_uint32]),
            "tm_cons3_download": (i, [vp, dp, dp, dp, dp, C.POINTER(C.c_uint64)]),
            "tm_cons3_last_ms": (i, [vp, C.POINTER(C.c_float), C.POINTER(i)]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _bound3 = True
    return L


class ConsumerOps3:
    """One tm_cons3_ctx: the vegetation + grazer + predator time step resident on the device."""

    def __init__(self, L, device=0):
        self._l = _lib3()
        self.L = L
        self._ctx = C.c_void_p()
        rc = self._l.tm_cons3_create(L, device, C.byref(self._ctx))
        if rc != 0:
            self._ctx = C.c_void_p()
            raise TmError("tm_cons3_create failed (%d): %s" % (rc, self._l.tm_cons3_last_error(None).decode()))

    def _ck(self, rc):
        if rc != 0:
            raise TmError("tm_cons3 call failed (%d): %s" % (rc, self._l.tm_cons3_last_error(self._ctx).decode()))

    def close(self):
        if self._ctx:
            self._l.tm_cons3_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setup(self, veg, graz, pred, cons, central, fr_graz, fr_pred):
        cen = np.ascontiguousarray(central, dtype=np.int32)
        self._ck(self._l.tm_cons3_setup(self._ctx, C.byref(veg), C.byref(graz), C.byref(pred), C.byref(cons),
                                        cen.ctypes.data_as(C.POINTER(C.c_int32)), len(cen), fr_graz, fr_pred))

    def set_fpe(self, terms, gauss_D, gauss_V, M1, M2):
        """Rietkerk_FPE flavour (tm_cons3_set_fpe); terms: (fke.FkeTerms * 2) for grazer and predator, or None for Rietkerk_FD."""
        dp = C.POINTER(C.c_double)
        if terms is None:
            self._ck(self._l.tm_cons3_set_fpe(self._ctx, None, None, None, 0.0, 0.0))
            return
        gD = [np.ascontiguousarray(x, dtype=np.float64) for x in gauss_D]
        gV = [np.ascontiguousarray(x, dtype=np.float64) for x in gauss_V]
        pD, pV = (dp * 2)(_dptr(gD[0]), _dptr(gD[1])), (dp * 2)(_dptr(gV[0]), _dptr(gV[1]))
        self._ck(self._l.tm_cons3_set_fpe(self._ctx, C.cast(terms, C.c_void_p), pD, pV, float(M1), float(M2)))

    def set_strict_means(self, on):
        self._ck(self._l.tm_cons3_set_strict_means(self._ctx, 1 if on else 0))

    def upload(self, fields, gamma=None):
        f = np.ascontiguousarray(fields, dtype=np.float64)
        assert f.size == 5 * self.L * self.L
        g = None if gamma is None else np.ascontiguousarray(gamma, dtype=np.float64)
        self._ck(self._l.tm_cons3_upload(self._ctx, _dptr(f), None if g is None else _dptr(g)))

    def run(self, seed, trial=0, step0=0, nsteps=1):
        self._ck(self._l.tm_cons3_run(self._ctx, seed, trial, step0, nsteps))

    def replay(self, u_veg, u_graz, u_pred, stride):
        us = [np.ascontiguousarray(x, dtype=np.float64) for x in (u_veg, u_graz, u_pred)]
        assert all(x.size == self.L * self.L * stride for x in us)
        self._ck(self._l.tm_cons3_replay(self._ctx, _dptr(us[0]), _dptr(us[1]), _dptr(us[2]), stride))

    def download(self):
        """-> (fields [5, L*L], gamma [2, L*L], v_eff [2, L*L, 2], means {<P>, <G>, <Pr>, #{P>0}, eff1, eff2}, overflow)."""
        n = self.L * self.L
        f, g, v, m, ovf = np.empty((5, n)), np.empty((2, n)), np.empty((2, n, 2)), np.empty(6), C.c_uint64()
        self._ck(self._l.tm_cons3_download(self._ctx, _dptr(f), _dptr(g), _dptr(v), _dptr(m), C.byref(ovf)))
        return f, g, v, m, ovf.value

    def last_ms(self):
        ms, n = C.c_float(), C.c_int()
        self._ck(self._l.tm_cons3_last_ms(self._ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value
