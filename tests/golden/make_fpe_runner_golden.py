"""Rietkerk_FPE SPB = 2 (vegetation + grazer) replicate-runner fixtures from the UNMODIFIED reference
(oracle/_ref/libtm_ref_fke.so: rietkerk_bjork_FKE.cpp compiled with -DSPB=2, CPU movement path):

  runner_fpe_2sp.npz   rietkerk_DorFPE_2D_MultiSp (F/:3697), one replicate under sequential replay of a seeded uniform sequence:
                       all four fields at full precision at t = 0 and after every one of the first steps, the number of uniforms
                       consumed, and the movement constants set_global_FKE_params derived (driver: order_2stoc_unity_rietkerk.cpp).

Run in the build container only (needs /root/reference):   python tests/golden/make_fpe_runner_golden.py
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
from make_runner2_golden import driver_par, frame, ptr, uniforms   # noqa: E402  (same constants and frame rule as the FD runner)


def main():
    os.chdir(tempfile.mkdtemp(prefix="tmgoldfpe_"))
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_fke.so"))
    assert ref.ref_fke_spb() == 2
    out = {}
    dtv1 = round(0.11817455 * 100.0) / 100.0                     # order_2stoc_unity_rietkerk.cpp:58, :62: dt = dt_analytical = dtv1
    for name, L, nsteps, seed, grazers in (("mixed", 24, 5, 41, True), ("no_grazers", 16, 3, 42, False)):
        par = np.concatenate([driver_par(dt=dtv1), [dtv1]])
        dt = par[21]
        init = frame(L, seed, grazers)
        n_u = 400000
        u = uniforms(seed, n_u)
        tmeas = np.array([0.0, dt * nsteps])
        frame_t = np.array([dt * k for k in range(1, nsteps + 1)] + [1e300])   # sentinel: the runner reads one past the end
        frames = np.zeros((32, 4, L * L)); ftimes = np.zeros(32); cons = C.c_uint64(); ov = C.c_uint64()
        k = ref.ref_fke_run(ptr(par), C.c_int(L), ptr(init), ptr(tmeas), C.c_int(len(tmeas)), ptr(frame_t), C.c_int(len(frame_t)),
                            ptr(u), C.c_uint64(n_u), ptr(frames), ptr(ftimes), C.c_int(32), C.byref(cons), C.byref(ov))
        assert k > 1 and ov.value == 0, (k, ov.value)
        derived = np.zeros(8)
        D = np.array([par[9], par[10], par[11], par[12]]); v = np.array([0.0, par[15]]); dta = np.array([0.0, par[23]])
        ref.ref_fke_set_params(ptr(D), ptr(v), ptr(dta), C.c_double(dt), C.c_double(par[22]), C.c_int(L), C.c_int(4), ptr(derived))
        print(name, "frames", k, "times", ftimes[:k], "consumed", cons.value, "derived", derived)
        out[name + "_par"] = par; out[name + "_L"] = np.array(L); out[name + "_seed"] = np.array([seed, n_u])
        out[name + "_init"] = init; out[name + "_frames"] = frames[:k].copy(); out[name + "_times"] = ftimes[:k].copy()
        out[name + "_consumed"] = np.array(cons.value); out[name + "_derived"] = derived
    R = int(driver_par()[20] / driver_par()[22])
    cen = np.zeros((4096, 2), dtype=np.int32)
    n = ref.ref_fke_central_sites(C.c_int(R), cen.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(4096))
    out["central_R%d" % R] = cen[:n].copy()
    np.savez_compressed(os.path.join(HERE, "runner_fpe_2sp.npz"), **out)


if __name__ == "__main__":
    main()
