"""Rietkerk_FPE SPB = 3 (vegetation + grazer + predator) replicate-runner fixtures from the UNMODIFIED reference
(oracle/_ref/libtm_ref_fke3.so: rietkerk_bjork_FKE.cpp compiled with -DSPB=3, CPU movement path):

  runner_fpe_3sp.npz   rietkerk_DorFPE_2D_MultiSp (F/:3697), one replicate under sequential replay: all five fields at full precision
                       at t = 0 and after the first steps, uniforms consumed, and the movement constants set_global_FKE_params derived
                       for grazer and predator.  Constants: the SPB = 3 runner vector of make_runner3_golden.py with
                       dt = dt_analytical = dtv1 (order_3stoc_unity_rietkerk.cpp) and dtAdv = {0, dtv1, dtv2}.

Run in the build container only (needs /root/reference):   python tests/golden/make_fpe_runner3_golden.py
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
from make_runner2_golden import ptr, uniforms        # noqa: E402
from make_runner3_golden import driver_par, frame    # noqa: E402


def main():
    os.chdir(tempfile.mkdtemp(prefix="tmgoldfpe3_"))
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_fke3.so"))
    assert ref.ref_fke3_spb() == 3
    out = {}
    dtv1, dtv2 = round(0.11817455 * 100.0) / 100.0, round(0.765270868 * 100.0) / 100.0
    for name, L, nsteps, seed, pred, graz in (("mixed", 22, 4, 71, True, True), ("no_predators", 16, 3, 72, False, True)):
        par = np.concatenate([driver_par(dt=dtv1), [dtv1, dtv2]])
        dt = par[29]
        init = frame(L, seed, pred, graz)
        n_u = 600000
        u = uniforms(seed, n_u)
        tmeas = np.array([0.0, dt * nsteps])
        frame_t = np.array([dt * k for k in range(1, nsteps + 1)] + [1e300])
        frames = np.zeros((32, 5, L * L)); ftimes = np.zeros(32); cons = C.c_uint64(); ov = C.c_uint64()
        k = ref.ref_fke3_run(ptr(par), C.c_int(L), ptr(init), ptr(tmeas), C.c_int(len(tmeas)), ptr(frame_t), C.c_int(len(frame_t)),
                             ptr(u), C.c_uint64(n_u), ptr(frames), ptr(ftimes), C.c_int(32), C.byref(cons), C.byref(ov))
        assert k > 1 and ov.value == 0, (k, ov.value)
        derived = np.zeros((2, 8))
        ref.ref_fke3_set_params(ptr(par), C.c_int(L), ptr(derived))
        print(name, "frames", k, "times", ftimes[:k], "consumed", cons.value, "derived", derived.tolist())
        out[name + "_par"] = par; out[name + "_L"] = np.array(L); out[name + "_seed"] = np.array([seed, n_u])
        out[name + "_init"] = init; out[name + "_frames"] = frames[:k].copy(); out[name + "_times"] = ftimes[:k].copy()
        out[name + "_consumed"] = np.array(cons.value); out[name + "_derived"] = derived
    np.savez_compressed(os.path.join(HERE, "runner_fpe_3sp.npz"), **out)


if __name__ == "__main__":
    main()
