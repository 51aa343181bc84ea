"""Output-tree fixture of the SPB = 2 replicate runner from the UNMODIFIED reference with ITS OWN CSV writers
(oracle/_ref/libtm_ref_riet2_files.so: rietkerk_Dornic_2D_MultiSp R/:3397, save_frame R/:2836, save_prelimframe R/:2720 and the two
wrappers R/:2886, R/:3029): two replicates under sequential replay write frames, gamma frames, PRELIM_TSERIES / PRELIM_MOVSERIES
checkpoints and PRELIM_AGGRAND tables.

  side_outputs_2sp.json.gz   {"files": {relative path: text}, "par": [...], "L": .., "r": .., "t_meas": [...], "frame_t": [...], "seed": ..,
                              "n_u": .., "init": [[...]], "consumed": ..}

Run in the build container only (needs /root/reference):   python tests/golden/make_side_outputs_golden.py
"""
import ctypes as C
import gzip
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)
from make_runner2_golden import driver_par, frame, ptr, uniforms   # noqa: E402


def main_fpe():
    """the Rietkerk_FPE flavour: side_outputs_fpe_2sp.json.gz from oracle/_ref/libtm_ref_fke_files.so (rietkerk_DorFPE_2D_MultiSp F/:3697)"""
    tmp = tempfile.mkdtemp(prefix="tmsidefpe_")
    os.chdir(tmp)
    root = os.path.join(tmp, "Rietkerk") + "/"
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_fke_files.so"))
    ref.reffkef_run.restype = C.c_uint64
    L, r, seed, n_u = 12, 2, 93, 400000
    dtv1 = round(0.11817455 * 100.0) / 100.0                     # order_2stoc_unity_rietkerk.cpp:58, :62
    par = np.concatenate([driver_par(dt=dtv1), [dtv1]])
    dt = par[21]
    init = np.stack([frame(L, seed + j) for j in range(r)])
    u = uniforms(seed, n_u)
    t_meas = np.array([0.0] + [dt * k for k in (1, 2, 3, 5, 7, 9, 12)])      # 8 bins: FPE checkpoints at indices 6, 7 (no 60 % mark)
    frame_t = np.array([dt * k for k in range(1, 14)] + [1e300])             # 14 entries: frames from frame_index 4 on (last ten), and t = 0
    a_tag = "%g-%g_dP_%g_Geq_%.5g" % (par[0], par[0], 10000.0, 4.8)
    os.makedirs(os.path.join(root, "Frames/Stochastic/2Sp", "SIDEFPE_" + a_tag))
    os.makedirs(os.path.join(root, "Prelims/Stochastic/2Sp", "SIDEFPE_" + a_tag, "TimeSeries"))
    os.makedirs(os.path.join(root, "Stochastic"))
    ov = C.c_uint64()
    cons = ref.reffkef_run(ptr(par), C.c_int(L), C.c_int(r), ptr(init), C.c_int(r), ptr(t_meas), C.c_int(len(t_meas)), ptr(frame_t),
                           C.c_int(len(frame_t)), ptr(u), C.c_uint64(n_u), root.encode(), b"SIDEFPE", C.c_double(4.8), C.byref(ov))
    assert ov.value == 0
    files = {}
    for dirpath, _, names in os.walk(root):
        for nm in names:
            p = os.path.join(dirpath, nm)
            files[os.path.relpath(p, root)] = open(p).read()
    print("FPE consumed", cons, "files", len(files))
    for k in sorted(files):
        print("  ", k, len(files[k]))
    out = {"files": files, "par": par.tolist(), "L": L, "r": r, "t_meas": t_meas.tolist(), "frame_t": frame_t[:-1].tolist(), "seed": seed,
           "n_u": n_u, "init": init.reshape(r, -1).tolist(), "consumed": int(cons), "Gstar": 4.8, "prefix": "SIDEFPE"}
    with gzip.open(os.path.join(HERE, "side_outputs_fpe_2sp.json.gz"), "wt") as f:
        json.dump(out, f)


def main_3sp():
    """SPB = 3: side_outputs_3sp.json.gz from oracle/_ref/libtm_ref_riet3_files.so"""
    import make_runner3_golden as G3
    tmp = tempfile.mkdtemp(prefix="tmside3_")
    os.chdir(tmp)
    root = os.path.join(tmp, "Rietkerk") + "/"
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_riet3_files.so"))
    ref.ref3f_run.restype = C.c_uint64
    L, r, seed, n_u = 10, 2, 95, 600000
    par = G3.driver_par()
    dt = par[29]
    init = np.stack([G3.frame(L, seed + j) for j in range(r)])
    u = uniforms(seed, n_u)
    t_meas = np.array([0.0] + [dt * k for k in (1, 2, 3, 5, 7, 9, 12)])
    frame_t = np.array([dt * k for k in (2, 5, 9, 12)] + [1e300])
    a_tag = "%g-%g_dP_%g_Geq_%.5g" % (par[0], par[0], 10000.0, 4.8)
    os.makedirs(os.path.join(root, "Frames/Stochastic/3Sp", "SIDE3_" + a_tag))
    os.makedirs(os.path.join(root, "Prelims/Stochastic/3Sp", "SIDE3_" + a_tag, "TimeSeries"))
    os.makedirs(os.path.join(root, "Stochastic"))
    ov = C.c_uint64()
    cons = ref.ref3f_run(ptr(par), C.c_int(L), C.c_int(r), ptr(init), C.c_int(r), ptr(t_meas), C.c_int(len(t_meas)), ptr(frame_t),
                         C.c_int(len(frame_t)), ptr(u), C.c_uint64(n_u), root.encode(), b"SIDE3", C.c_double(4.8), C.byref(ov))
    assert ov.value == 0
    files = {}
    for dirpath, _, names in os.walk(root):
        for nm in names:
            p = os.path.join(dirpath, nm)
            files[os.path.relpath(p, root)] = open(p).read()
    print("3Sp consumed", cons, "files", len(files))
    out = {"files": files, "par": par.tolist(), "L": L, "r": r, "t_meas": t_meas.tolist(), "frame_t": frame_t[:-1].tolist(), "seed": seed,
           "n_u": n_u, "init": init.reshape(r, -1).tolist(), "consumed": int(cons), "Gstar": 4.8, "prefix": "SIDE3"}
    with gzip.open(os.path.join(HERE, "side_outputs_3sp.json.gz"), "wt") as f:
        json.dump(out, f)


def main_fpe3():
    """Rietkerk_FPE, SPB = 3: side_outputs_fpe_3sp.json.gz from oracle/_ref/libtm_ref_fke3_files.so"""
    import make_runner3_golden as G3
    tmp = tempfile.mkdtemp(prefix="tmsidefpe3_")
    os.chdir(tmp)
    root = os.path.join(tmp, "Rietkerk") + "/"
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_fke3_files.so"))
    ref.reffke3f_run.restype = C.c_uint64
    L, r, seed, n_u = 12, 2, 97, 600000
    dtv1, dtv2 = round(0.11817455 * 100.0) / 100.0, round(0.765270868 * 100.0) / 100.0
    par = np.concatenate([G3.driver_par(dt=dtv1), [dtv1, dtv2]])
    dt = par[29]
    init = np.stack([G3.frame(L, seed + j) for j in range(r)])
    u = uniforms(seed, n_u)
    t_meas = np.array([0.0] + [dt * k for k in (1, 2, 3, 5, 7, 9, 12)])
    frame_t = np.array([dt * k for k in range(1, 14)] + [1e300])
    a_tag = "%g-%g_dP_%g_Geq_%.5g" % (par[0], par[0], 10000.0, 4.8)
    os.makedirs(os.path.join(root, "Frames/Stochastic/3Sp", "SIDEFPE3_" + a_tag))
    os.makedirs(os.path.join(root, "Prelims/Stochastic/3Sp", "SIDEFPE3_" + a_tag, "TimeSeries"))
    os.makedirs(os.path.join(root, "Stochastic"))
    derived = np.zeros((2, 8))
    ref.reffke3f_set_params(ptr(par), C.c_int(L), ptr(derived))
    ov = C.c_uint64()
    cons = ref.reffke3f_run(ptr(par), C.c_int(L), C.c_int(r), ptr(init), C.c_int(r), ptr(t_meas), C.c_int(len(t_meas)), ptr(frame_t),
                            C.c_int(len(frame_t)), ptr(u), C.c_uint64(n_u), root.encode(), b"SIDEFPE3", C.c_double(4.8), C.byref(ov))
    assert ov.value == 0
    files = {}
    for dirpath, _, names in os.walk(root):
        for nm in names:
            p = os.path.join(dirpath, nm)
            files[os.path.relpath(p, root)] = open(p).read()
    print("FPE 3Sp consumed", cons, "files", len(files))
    out = {"files": files, "par": par.tolist(), "L": L, "r": r, "t_meas": t_meas.tolist(), "frame_t": frame_t[:-1].tolist(), "seed": seed,
           "n_u": n_u, "init": init.reshape(r, -1).tolist(), "consumed": int(cons), "Gstar": 4.8, "prefix": "SIDEFPE3",
           "derived": derived.tolist()}
    with gzip.open(os.path.join(HERE, "side_outputs_fpe_3sp.json.gz"), "wt") as f:
        json.dump(out, f)


def main():
    tmp = tempfile.mkdtemp(prefix="tmside_")
    os.chdir(tmp)
    root = os.path.join(tmp, "Rietkerk") + "/"
    for d in ("Frames/Stochastic/2Sp", "Prelims/Stochastic/2Sp", "Stochastic"):
        os.makedirs(os.path.join(root, d))
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtm_ref_riet2_files.so"))
    ref.ref2f_run.restype = C.c_uint64
    L, r, seed, n_u = 10, 2, 91, 400000
    par = driver_par()
    dt = par[21]
    init = np.stack([frame(L, seed + j) for j in range(r)])
    u = uniforms(seed, n_u)
    t_meas = np.array([0.0] + [dt * k for k in (1, 2, 3, 5, 7, 9, 12)])      # 8 bins: checkpoints at indices 4, 6, 7, 7, 7
    frame_t = np.array([dt * k for k in (2, 5, 9, 12)] + [1e300])
    # the wrappers create their folders as they go in the reference's drivers (recursive_dir_create in main()); make them here
    a_tag = "%g-%g_dP_%g_Geq_%.5g" % (par[0], par[0], 10000.0, 4.8)
    os.makedirs(os.path.join(root, "Frames/Stochastic/2Sp", "SIDE_" + a_tag))
    os.makedirs(os.path.join(root, "Prelims/Stochastic/2Sp", "SIDE_" + a_tag, "TimeSeries"))
    ov = C.c_uint64()
    cons = ref.ref2f_run(ptr(par), C.c_int(L), C.c_int(r), ptr(init), C.c_int(r), ptr(t_meas), C.c_int(len(t_meas)), ptr(frame_t),
                         C.c_int(len(frame_t)), ptr(u), C.c_uint64(n_u), root.encode(), b"SIDE", C.c_double(4.8), C.byref(ov))
    assert ov.value == 0
    files = {}
    for dirpath, _, names in os.walk(root):
        for nm in names:
            p = os.path.join(dirpath, nm)
            files[os.path.relpath(p, root)] = open(p).read()
    print("consumed", cons, "files", len(files))
    for k in sorted(files):
        print("  ", k, len(files[k]))
    out = {"files": files, "par": par.tolist(), "L": L, "r": r, "t_meas": t_meas.tolist(), "frame_t": frame_t[:-1].tolist(), "seed": seed,
           "n_u": n_u, "init": init.reshape(r, -1).tolist(), "consumed": int(cons), "Gstar": 4.8, "prefix": "SIDE"}
    with gzip.open(os.path.join(HERE, "side_outputs_2sp.json.gz"), "wt") as f:
        json.dump(out, f)


if __name__ == "__main__":
    if "--fpe3-only" in sys.argv:
        main_fpe3()
        sys.exit(0)
    if "--3sp-only" in sys.argv:
        main_3sp()
        sys.exit(0)
    if "--fpe-only" not in sys.argv:
        main()
        main_3sp()
        main_fpe3()
    main_fpe()
