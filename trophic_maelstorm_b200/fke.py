"""ctypes view of include/tm_fke.h (no numerics live here): the FKE movement step of the consumer species."""
import ctypes as C

import numpy as np

from .lattice import TmError, lib


class FkeTerms(C.Structure):
    """struct tm_fke_terms."""
    _fields_ = [("bound_D", C.c_int32), ("v_row_lo", C.c_int32), ("v_row_hi", C.c_int32), ("v_col_lo", C.c_int32),
                ("v_col_hi", C.c_int32), ("n_gauss_D", C.c_int32), ("n_gauss_V", C.c_int32), ("pad_", C.c_int32),
                ("mu_vel_prefactor", C.c_double), ("epsilon", C.c_double)]


FKE_ABI_SYMBOLS = [
    "tm_fke_create", "tm_fke_destroy", "tm_fke_last_error", "tm_fke_set_terms", "tm_fke_set_terms_device", "tm_fke_move_host", "tm_fke_move_device",
    "tm_fke_sync", "tm_fke_cuda_stream", "tm_fke_last_ms", "tm_fke_last_displacement",
]

_bound = False


def _lib():
    global _bound
    L = lib()
    if not _bound:
        vp, dp, i = C.c_void_p, C.POINTER(C.c_double), C.c_int
        sig = {
            "tm_fke_create": (i, [i, i, i, C.POINTER(vp)]),
            "tm_fke_destroy": (i, [vp]),
            "tm_fke_last_error": (C.c_char_p, [vp]),
            "tm_fke_set_terms": (i, [vp, i, C.POINTER(FkeTerms), dp, dp]),
            "tm_fke_set_terms_device": (i, [vp, i, C.POINTER(FkeTerms), vp, vp]),
            "tm_fke_move_host": (i, [vp, dp, dp, dp, dp]),
            "tm_fke_move_device": (i, [vp, vp, vp, vp, vp]),
            "tm_fke_sync": (i, [vp]),
            "tm_fke_cuda_stream": (vp, [vp]),
            "tm_fke_last_ms": (i, [vp, C.POINTER(C.c_float), C.POINTER(i)]),
            "tm_fke_last_displacement": (i, [vp, i, C.POINTER(C.c_int32)]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _bound = True
    return L


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class FkeMover:
    """One tm_fke_ctx: the movement step of n_species consumer species on an L x L periodic lattice."""

    def __init__(self, L, n_species=1, device=0):
        self._l = _lib()
        self.L, self.n_species = L, n_species
        self._ctx = C.c_void_p()
        rc = self._l.tm_fke_create(L, n_species, device, C.byref(self._ctx))
        if rc != 0:
            self._ctx = C.c_void_p()
            raise TmError("tm_fke_create failed (%d): %s" % (rc, self._l.tm_fke_last_error(None).decode()))

    def _ck(self, rc):
        if rc != 0:
            raise TmError("tm_fke call failed (%d): %s" % (rc, self._l.tm_fke_last_error(self._ctx).decode()))

    def close(self):
        if self._ctx:
            self._l.tm_fke_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_terms(self, s, bound_D, rows, cols, prefactor, epsilon, gauss_D, gauss_V):
        """rows / cols: (lo, hi) of the advection window along each axis."""
        gD = np.ascontiguousarray(gauss_D, dtype=np.float64)
        gV = np.ascontiguousarray(gauss_V, dtype=np.float64)
        t = FkeTerms(bound_D, rows[0], rows[1], cols[0], cols[1], gD.size, gV.size, 0, prefactor, epsilon)
        self._ck(self._l.tm_fke_set_terms(self._ctx, s, C.byref(t), _dptr(gD), _dptr(gV)))

    def move(self, rho, gamma, veff):
        """Host buffers [n_species][L*L] (veff [n_species][L*L][2]) -> moved rho, same shape as rho."""
        n = self.n_species * self.L * self.L
        rho, gamma, veff = (np.ascontiguousarray(a, dtype=np.float64) for a in (rho, gamma, veff))
        assert rho.size == n and gamma.size == n and veff.size == 2 * n
        out = np.empty_like(rho)
        self._ck(self._l.tm_fke_move_host(self._ctx, _dptr(rho), _dptr(gamma), _dptr(veff), _dptr(out)))
        return out

    def move_device(self, d_rho, d_gamma, d_veff, d_out):
        """Raw device pointers (ints); asynchronous on the context's stream."""
        self._ck(self._l.tm_fke_move_device(self._ctx, C.c_void_p(d_rho), C.c_void_p(d_gamma), C.c_void_p(d_veff), C.c_void_p(d_out)))

    def sync(self):
        self._ck(self._l.tm_fke_sync(self._ctx))

    def cuda_stream(self):
        return self._l.tm_fke_cuda_stream(self._ctx)

    def last_ms(self):
        ms, n = C.c_float(), C.c_int()
        self._ck(self._l.tm_fke_last_ms(self._ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_displacement(self, s=0):
        out = (C.c_int32 * 2)()
        self._ck(self._l.tm_fke_last_displacement(self._ctx, s, out))
        return out[0], out[1]
