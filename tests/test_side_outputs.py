"""CSV side outputs of the multi-species replicate runners (VERDICT r1 missing #5): FRAME_* / GAMMA_* snapshots, the
movement averages Rho_Mov, the PRELIM_TSERIES_* / PRELIM_MOVSERIES_* checkpoints and the PRELIM_AGGRAND_* running tables.

The fixture tests/golden/side_outputs_2sp.json.gz is the output tree the UNMODIFIED reference runner (rietkerk_Dornic_2D_MultiSp, R/:3397,
-DSPB=2) wrote with ITS OWN writers (save_frame R/:2836, save_prelimframe R/:2720, the wrappers R/:2886, R/:3029) under sequential
replay (tests/golden/make_side_outputs_golden.py).  Here the host layer's runner (tm_b200::spb2::rietkerk_Dornic_2D_MultiSp: its
measurement schedule, running statistics and writers, tm_host_side.h) runs over its test seam with the CPU oracle as the Stepper
-- no device -- and must write the same tree, byte for byte, up to the reference's two random file-name tokens (jID, ID: constant 0
in the host layer).  No GPU needed: this pins the bookkeeping; the device step it drives in production is pinned elsewhere
(test_consumer_step2.py, test_zz_*.py)."""
import ctypes as C
import gzip
import json
import os
import re

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
dp = C.POINTER(C.c_double)
GOLD = np.load(os.path.join(HERE, "golden", "consumer_ops.npz"))


@pytest.fixture(scope="module")
def host():
    C.CDLL(os.path.join(ROOT, "trophic_maelstorm_b200", "libtm_lattice.so"), mode=C.RTLD_GLOBAL)
    return C.CDLL(os.path.join(ROOT, "trophic_maelstorm_b200", "libtm_host.so"))


def _normalise(path):
    """the reference's random tokens -> the host layer's constant 0"""
    path = re.sub(r"_jID_\d+_", "_jID_0_", path)
    return re.sub(r"_ID_[0-9.e+-]+_DP_", "_ID_0_DP_", path)


def _loop_iterations(t_end, dt):
    t, k = 0.0, 0
    while t < t_end + dt:                      # R/:3648
        t += dt
        k += 1
    return k


class OracleStepper2:
    """The SPB = 2 oracle step (tmo_step_2sp, pinned on the reference's frames in test_consumer_oracle.py) behind the three Stepper
    calls, consuming ONE sequential replay stream exactly as the reference's runner does: thrID before replicate 0 (R/:3634), every
    time step of the `while (t < t_max + dt)` loop -- the ones after the last bookkeeping event included --, one draw for the running
    table's file name after every replicate (R/:4343)."""

    def __init__(self, L, par, u, n_loop):
        self.L, self.n_loop = L, n_loop
        self.pv, self.pg, self.pc, R, self.fr = O.runner2_params(L, par)
        self.central = np.ascontiguousarray(GOLD["central_R%d" % R], dtype=np.int32)
        self.u = u
        self.rng = O.make_rng(O.SRC_SEQ, u=u, pos=0)
        self.f = np.zeros((4, L * L))
        self.gamma = np.zeros(L * L)
        self.veff = np.zeros((L * L, 2))
        self.replicate = -1
        self.steps_done = 0
        self.calls = []

    def _one_step(self, lib):
        return lib.tmo_step_2sp_means(C.byref(self.pv), C.byref(self.pg), C.byref(self.pc),
                                      self.central.ctypes.data_as(C.POINTER(C.c_int32)), len(self.central), C.c_double(self.fr),
                                      C.byref(self.rng), O.ptr(self.f), O.ptr(self.gamma), O.ptr(self.veff), 1, None)

    def _step(self, n):
        lib = O.oracle()
        for _ in range(n):
            assert self._one_step(lib) == 0
        self.steps_done += n

    def finish_replicate(self):
        if self.replicate >= 0:
            self._step(self.n_loop - self.steps_done)      # the loop's steps after the last measurement / frame
            self.rng.pos += 1                              # ran_jid, R/:4343

    def upload(self, frame):
        self.finish_replicate()
        self.replicate += 1
        if self.replicate == 0:
            self.rng.pos += 1                              # thrID, R/:3634
        n = 4 * self.L * self.L
        self.f[:] = np.ctypeslib.as_array(frame, shape=(n,)).reshape(4, -1)
        self.gamma[:] = 0.0                                # R/:3523
        self.veff[:] = 0.0                                 # R/:3526
        self.steps_done = 0
        self.rng.norm_saved_available = 0                  # every replicate has its own `norm` object (R/:3628)

    def run(self, replicate, step0, nsteps):
        assert replicate == self.replicate and step0 == self.steps_done
        self.calls.append((replicate, step0, nsteps))
        self._step(nsteps)

    def download(self, frame, gamma, veff):
        n = self.L * self.L
        np.ctypeslib.as_array(frame, shape=(4 * n,))[:] = self.f.reshape(-1)
        if gamma:
            np.ctypeslib.as_array(gamma, shape=(n,))[:] = self.gamma
        if veff:
            np.ctypeslib.as_array(veff, shape=(2 * n,))[:] = self.veff.reshape(-1)


class OracleStepperFpe2(OracleStepper2):
    """the same for the Rietkerk_FPE runner of the SPB = 2 model (tmo_step_fpe_2sp, pinned in test_fpe_runner_oracle.py); the movement
    constants are the ones set_global_FKE_params derives (tmo_fke_params, pinned on the reference in test_fke_oracle.py)"""

    def __init__(self, L, par, u, n_loop):
        OracleStepper2.__init__(self, L, par[:23], u, n_loop)
        d = O.fke_params(par[10], par[15], par[23], par[22])
        derived = [d.beta_vel, d.sigma_D, d.sigma_vD, d.mu_vel_prefactor, d.bound_D, d.bound_vD_first, d.bound_vD_second, 1.0e-10]
        self.pv, self.pc, self.M1, self.terms, self.gD, self.gV, R, self.fr = O.fpe_runner_params(L, par, derived)

    def _one_step(self, lib):
        return lib.tmo_step_fpe_2sp(C.byref(self.pv), C.byref(self.pc), C.c_double(self.M1), C.byref(self.terms), O.ptr(self.gD),
                                    O.ptr(self.gV), self.central.ctypes.data_as(C.POINTER(C.c_int32)), len(self.central),
                                    C.c_double(self.fr), C.byref(self.rng), O.ptr(self.f), O.ptr(self.gamma), O.ptr(self.veff), None)


class OracleStepper3(OracleStepper2):
    """SPB = 3: tmo_step_3sp (pinned in test_consumer3_oracle.py) behind the same three calls; gamma [2][n], veff [2][n][2]"""

    def __init__(self, L, par, u, n_loop):
        self.L, self.n_loop = L, n_loop
        self.pv, self.pg1, self.pg2, self.pc, R, self.fr1, self.fr2 = O.runner3_params(L, par)
        G3 = np.load(os.path.join(HERE, "golden", "runner_3sp.npz"))
        self.central = np.ascontiguousarray(G3["central_R%d" % R], dtype=np.int32)
        self.u = u
        self.rng = O.make_rng(O.SRC_SEQ, u=u, pos=0)
        self.f = np.zeros((5, L * L))
        self.gamma = np.zeros((2, L * L))
        self.veff = np.zeros((2, L * L, 2))
        self.replicate, self.steps_done, self.calls = -1, 0, []

    def _one_step(self, lib):
        return lib.tmo_step_3sp(C.byref(self.pv), C.byref(self.pg1), C.byref(self.pg2), C.byref(self.pc),
                                self.central.ctypes.data_as(C.POINTER(C.c_int32)), len(self.central), C.c_double(self.fr1),
                                C.c_double(self.fr2), C.byref(self.rng), O.ptr(self.f), O.ptr(self.gamma), O.ptr(self.veff), 1, None)

    def upload(self, frame):
        self.finish_replicate()
        self.replicate += 1
        if self.replicate == 0:
            self.rng.pos += 1
        n = 5 * self.L * self.L
        self.f[:] = np.ctypeslib.as_array(frame, shape=(n,)).reshape(5, -1)
        self.gamma[:] = 0.0
        self.veff[:] = 0.0
        self.steps_done = 0
        self.rng.norm_saved_available = 0

    def download(self, frame, gamma, veff):
        n = self.L * self.L
        np.ctypeslib.as_array(frame, shape=(5 * n,))[:] = self.f.reshape(-1)
        if gamma:
            np.ctypeslib.as_array(gamma, shape=(2 * n,))[:] = self.gamma.reshape(-1)
        if veff:
            np.ctypeslib.as_array(veff, shape=(4 * n,))[:] = self.veff.reshape(-1)


class OracleStepperFpe3(OracleStepper3):
    """Rietkerk_FPE, SPB = 3: tmo_step_fpe_3sp (pinned in test_fpe_runner_oracle.py); the advection vectors are state there"""

    def __init__(self, L, par, u, n_loop, derived):
        OracleStepper3.__init__(self, L, par[:31], u, n_loop)
        self.pv, self.pc, self.M, self.v, self.terms, self.gD, self.gV, R, self.fr1, self.fr2 = O.fpe_runner3_params(L, par, derived)
        self.pD = (dp * 2)(O.ptr(self.gD[0]), O.ptr(self.gD[1]))
        self.pV = (dp * 2)(O.ptr(self.gV[0]), O.ptr(self.gV[1]))

    def _one_step(self, lib):
        return lib.tmo_step_fpe_3sp(C.byref(self.pv), C.byref(self.pc), O.ptr(self.M), O.ptr(self.v), self.terms, self.pD, self.pV,
                                    self.central.ctypes.data_as(C.POINTER(C.c_int32)), len(self.central), C.c_double(self.fr1),
                                    C.c_double(self.fr2), C.byref(self.rng), O.ptr(self.f), O.ptr(self.gamma), O.ptr(self.veff), 1, None)


UP = C.CFUNCTYPE(None, dp)
RUN = C.CFUNCTYPE(None, C.c_uint32, C.c_uint32, C.c_int)
DOWN = C.CFUNCTYPE(None, dp, dp, dp)


def _side_run(host, spb, L, par, r, init, t_meas, frame_t, root, prefix, Gstar, st, width):
    host.tmh_side_run.restype = C.c_int
    host.tmh_side_run.argtypes = [C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_char_p, C.c_char_p,
                                  C.c_double, C.c_double, UP, RUN, DOWN, dp, C.c_int, C.c_int]
    rows = np.zeros((len(t_meas) + 4, width))
    cb = (UP(st.upload), RUN(st.run), DOWN(st.download))
    n = host.tmh_side_run(spb, L, O.ptr(par), r, O.ptr(init), len(init), O.ptr(t_meas), len(t_meas), O.ptr(frame_t), len(frame_t),
                          root.encode(), prefix.encode(), Gstar, -1.0, cb[0], cb[1], cb[2], O.ptr(rows), len(rows), 1)
    assert n == len(t_meas), n
    return rows[:n]


def _tree(root):
    files = {}
    for dirpath, _, names in os.walk(root):
        for nm in names:
            p = os.path.join(dirpath, nm)
            files[os.path.relpath(p, root)] = open(p).read()
    return files


def test_spb2_output_tree_equals_the_reference(host, tmp_path):
    with gzip.open(os.path.join(HERE, "golden", "side_outputs_2sp.json.gz"), "rt") as f:
        g = json.load(f)
    L, r = g["L"], g["r"]
    par = np.array(g["par"])
    t_meas = np.array(g["t_meas"])
    frame_t = np.array(g["frame_t"] + [1e300])             # the sentinel the fixture run ended its frame schedule with
    init = np.array(g["init"]).reshape(r, 4, L * L)
    u = np.random.RandomState(g["seed"]).random_sample(g["n_u"])
    st = OracleStepper2(L, par, u, _loop_iterations(t_meas[-1], par[21]))
    root = str(tmp_path) + "/"
    rows = _side_run(host, 2, L, par, r, init, t_meas, frame_t, root, g["prefix"], g["Gstar"], st, 14)
    st.finish_replicate()
    assert st.rng.overflow == 0
    assert st.rng.pos == g["consumed"]                      # every uniform of the reference run is accounted for

    want = {_normalise(k): v for k, v in g["files"].items()}
    assert len(want) == len(g["files"])
    got = _tree(root)
    assert sorted(got) == sorted(want)
    for k in sorted(want):
        assert got[k] == want[k], k
    # steps between two bookkeeping events go to the device as one batch
    assert max(n for _, _, n in st.calls) > 1
    # the returned rows are the last replicate's running table (PRELIM_AGGRAND_*_R_<r-1>) at full precision
    agg = [k for k in want if "AGGRAND" in k and k.endswith("_R_%d.csv" % (r - 1))]
    assert len(agg) == 1
    tab = np.array([[float(x) for x in line.split(",")[:-1]] for line in want[agg[0]].splitlines()[1:] if line.strip()])
    assert tab.shape == (len(t_meas), 3 + 11)
    np.testing.assert_allclose(rows[:, 3:], tab[:, 3:], rtol=2e-5, atol=1e-300)
    assert np.array_equal(rows[:, 0], np.full(len(t_meas), par[0])) and np.array_equal(rows[:, 1], np.full(len(t_meas), r))


def test_fpe_spb2_output_tree_equals_the_reference(host, tmp_path):
    """rietkerk_DorFPE_2D_MultiSp (F/:3697): the Rietkerk_FPE flavour of the bookkeeping -- checkpoints at the 85 / 90 / 95 / 100 % bins
    only (F/:3032), frames by its own rule (F/:2886: here t = 0 and the last ten of the schedule) -- against the tree the unmodified
    reference wrote (oracle/ref_harness_fke_files.cpp)."""
    with gzip.open(os.path.join(HERE, "golden", "side_outputs_fpe_2sp.json.gz"), "rt") as f:
        g = json.load(f)
    L, r = g["L"], g["r"]
    par = np.array(g["par"])
    t_meas = np.array(g["t_meas"])
    frame_t = np.array(g["frame_t"] + [1e300])
    init = np.array(g["init"]).reshape(r, 4, L * L)
    u = np.random.RandomState(g["seed"]).random_sample(g["n_u"])
    st = OracleStepperFpe2(L, par, u, _loop_iterations(t_meas[-1], par[21]))
    root = str(tmp_path) + "/"
    rows = _side_run(host, 12, L, par, r, init, t_meas, frame_t, root, g["prefix"], g["Gstar"], st, 14)
    st.finish_replicate()
    assert st.rng.overflow == 0 and st.rng.pos == g["consumed"]
    want = {_normalise(k): v for k, v in g["files"].items()}
    got = _tree(root)
    assert sorted(got) == sorted(want)
    for k in sorted(want):
        if "MOVSERIES" in k:
            # Rietkerk_FPE's save_prelimframe walks 3 + 4 SpB columns of every table (F/:2758-2763): two doubles past the end of a
            # movement row, heap garbage in the reference's file; the host layer writes the same number of columns with zeros there
            w = ["" if i == 0 else ",".join(ln.split(",")[:12] + ["0", "0", ""]) for i, ln in enumerate(want[k].splitlines())]
            gl = got[k].splitlines()
            assert gl[0] == want[k].splitlines()[0] and gl[1:] == w[1:], k
        else:
            assert got[k] == want[k], k
    assert not any("_T_0.6_" in k and "PRELIM" in k for k in got)        # index 4 = int(0.6 * 8): a checkpoint of Rietkerk_FD only
    assert np.isfinite(rows).all()


def test_spb3_output_tree_equals_the_reference(host, tmp_path):
    """the same for the vegetation + grazer + predator runner (R/:3397 with -DSPB=3): five-field frames, both perception fields, the
    3Sp headers (rietkerk_bjork_constants_3Sp.h:6-17), 18-column tables, twelve movement columns"""
    with gzip.open(os.path.join(HERE, "golden", "side_outputs_3sp.json.gz"), "rt") as f:
        g = json.load(f)
    L, r = g["L"], g["r"]
    par = np.array(g["par"])
    t_meas = np.array(g["t_meas"])
    frame_t = np.array(g["frame_t"] + [1e300])
    init = np.array(g["init"]).reshape(r, 5, L * L)
    u = np.random.RandomState(g["seed"]).random_sample(g["n_u"])
    st = OracleStepper3(L, par, u, _loop_iterations(t_meas[-1], par[29]))
    root = str(tmp_path) + "/"
    rows = _side_run(host, 3, L, par, r, init, t_meas, frame_t, root, g["prefix"], g["Gstar"], st, 18)
    st.finish_replicate()
    assert st.rng.overflow == 0 and st.rng.pos == g["consumed"]
    want = {_normalise(k): v for k, v in g["files"].items()}
    got = _tree(root)
    assert sorted(got) == sorted(want)
    for k in sorted(want):
        assert got[k] == want[k], k
    assert any("GAM[Pr(x; t)]" in v for k, v in got.items() if "GAMMA_" in k) and np.isfinite(rows).all()


def test_fpe_spb3_output_tree_equals_the_reference(host, tmp_path):
    """rietkerk_DorFPE_2D_MultiSp with -DSPB=3 (F/:3697): the FPE flavour of the bookkeeping with five fields, two perception fields and
    the per-site advection directions in the movement tables (fifteen columns walked by F/'s writer: two past the row)"""
    with gzip.open(os.path.join(HERE, "golden", "side_outputs_fpe_3sp.json.gz"), "rt") as f:
        g = json.load(f)
    L, r = g["L"], g["r"]
    par = np.array(g["par"])
    t_meas = np.array(g["t_meas"])
    frame_t = np.array(g["frame_t"] + [1e300])
    init = np.array(g["init"]).reshape(r, 5, L * L)
    u = np.random.RandomState(g["seed"]).random_sample(g["n_u"])
    st = OracleStepperFpe3(L, par, u, _loop_iterations(t_meas[-1], par[29]), np.array(g["derived"]))
    root = str(tmp_path) + "/"
    rows = _side_run(host, 13, L, par, r, init, t_meas, frame_t, root, g["prefix"], g["Gstar"], st, 18)
    st.finish_replicate()
    assert st.rng.overflow == 0 and st.rng.pos == g["consumed"]
    want = {_normalise(k): v for k, v in g["files"].items()}
    got = _tree(root)
    assert sorted(got) == sorted(want)
    for k in sorted(want):
        if "MOVSERIES" in k:    # F/:2758-2763 walks 3 + 4 SpB = 15 columns of a 13-column movement row: zeros here for the two beyond
            w = [",".join(ln.split(",")[:16] + ["0", "0", ""]) for ln in want[k].splitlines()]
            gl = got[k].splitlines()
            assert gl[0] == want[k].splitlines()[0] and gl[1:] == w[1:], k
        else:
            assert got[k] == want[k], k
    assert np.isfinite(rows).all()
    mov = [v for k, v in got.items() if "MOVSERIES" in k][0].splitlines()
    assert any(float(x) != 0 for x in mov[2].split(",")[10:12])           # <v_x>, <v_y> of the grazer: the drawn directions are reported


def test_side_outputs_off(host, tmp_path):
    """no frame schedule and write_prelims off: nothing is written, no movement fields are fetched"""
    with gzip.open(os.path.join(HERE, "golden", "side_outputs_2sp.json.gz"), "rt") as f:
        g = json.load(f)
    L, r = g["L"], 1
    par = np.array(g["par"])
    t_meas = np.array(g["t_meas"])
    init = np.array(g["init"]).reshape(g["r"], 4, L * L)[:1]
    u = np.random.RandomState(g["seed"]).random_sample(g["n_u"])
    st = OracleStepper2(L, par, u, _loop_iterations(t_meas[-1], par[21]))
    fetched = []
    inner = st.download
    st.download = lambda f, gm, ve: (fetched.append((bool(gm), bool(ve))), inner(f, gm, ve))[1]
    host.tmh_side_run.restype = C.c_int
    rows = np.zeros((len(t_meas) + 4, 14))
    cb = (UP(st.upload), RUN(st.run), DOWN(st.download))
    host.tmh_side_run.argtypes = [C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_char_p, C.c_char_p,
                                  C.c_double, C.c_double, UP, RUN, DOWN, dp, C.c_int, C.c_int]
    n = host.tmh_side_run(2, L, O.ptr(par), r, O.ptr(init), 1, O.ptr(t_meas), len(t_meas), None, 0, (str(tmp_path) + "/").encode(),
                          b"OFF", -1.0, -1.0, cb[0], cb[1], cb[2], O.ptr(rows), len(rows), 0)
    assert n == len(t_meas)
    assert _tree(str(tmp_path)) == {}
    assert fetched and not any(gm or ve for gm, ve in fetched)
