"""Strict means of the multi-species steps: with tm_cons_step2_set_strict_means / tm_cons3_set_strict_means the
spatial means <P>, <G> (, <Pr>) come from mean_of_vector's own left-to-right sums (R/:1211-1219; cons_seqsum_kernel, one thread per
field) instead of the device reduction tree.  The step then equals the reference-pinned oracle BIT FOR BIT FREE-RUNNING -- the oracle
computes its own means, nothing is handed over -- over several steps, where the default (tree) mode is bit-equal only given the
product's means and otherwise a few ulp off in the perception field.  CPU: the kernel source on host threads; GPU: the C ABI."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import test_consumer_step2 as S2
import test_zz_consumer_step3 as S3
from test_consumer_step2 import emu   # noqa: F401  (fixture)


@pytest.fixture()
def strict_emu(emu):   # noqa: F811
    emu.emu_set_strict_means(1)
    yield emu
    emu.emu_set_strict_means(0)


@pytest.mark.parametrize("L,nsteps", [(20, 4), (48, 3)])
def test_emulated_step2_strict_means_free_running(strict_emu, L, nsteps):
    par, pv, pg, pc, central, fr, f = S2.setup(L, 5 + L)
    gamma = np.zeros(L * L)
    wf, wg, wv = S2.oracle_steps(pv, pg, pc, central, fr, f, gamma, nsteps, seed=77, trial=3, step0=10)
    gf, gg, gv, dyn, ovf = S2.emu_step2(strict_emu, L, par, central, fr, f, gamma, nsteps=nsteps, seed=77, trial=3, step0=10)
    assert ovf == 0
    assert np.array_equal(gf, wf.reshape(gf.shape)) and np.array_equal(gg, wg) and np.array_equal(gv, wv)


def test_emulated_step2_default_means_are_the_tree(emu):   # noqa: F811
    """the switch is what makes the difference: in the default mode the means differ from the sequential sums in the last bits"""
    L = 48
    par, pv, pg, pc, central, fr, f = S2.setup(L, 5 + L)
    tree = S2.emu_step2(emu, L, par, central, fr, f, np.zeros(L * L), nsteps=1, seed=1)[3]
    emu.emu_set_strict_means(1)
    try:
        strict = S2.emu_step2(emu, L, par, central, fr, f, np.zeros(L * L), nsteps=1, seed=1)[3]
    finally:
        emu.emu_set_strict_means(0)
    assert strict[0] == O.mean(f[0]) and strict[1] == O.mean(f[1]) and strict[2] == O.occupied(f[0])
    assert strict[2] == tree[2] and strict[3] == tree[3]
    assert abs(tree[0] - strict[0]) <= 1e-13 * strict[0] and abs(tree[1] - strict[1]) <= 1e-13 * strict[1]


def test_emulated_step3_strict_means_free_running(strict_emu):
    L, nsteps = 16, 3
    par, pv, pg1, pg2, pc, central, fr1, fr2, f = S3.setup(L, 7)
    fo, go = f.copy(), np.zeros((2, L * L))
    fp, gp = f.copy(), np.zeros((2, L * L))
    for s in range(nsteps):
        fp, gp, vp, dyn, ovf = S3.emu_step(strict_emu, L, par, central, fr1, fr2, fp, gp, s, 99)
        assert ovf == 0 and [dyn[k] for k in range(3)] == [O.mean(fo[k]) for k in range(3)]
        vo = O.step_3sp(pv, pg1, pg2, pc, central, fr1, fr2, O.make_rng(O.SRC_PHILOX, seed=99, trial=0, step=s), fo, go)
        assert np.array_equal(fp, fo) and np.array_equal(gp, go) and np.array_equal(vp, vo), "step %d" % s


@pytest.mark.gpu
@pytest.mark.parametrize("L,nsteps", [(20, 4), (150, 3), (512, 2)])
def test_gpu_step2_strict_means_free_running(L, nsteps):
    import trophic_maelstorm_b200 as tm
    par, pv, pg, pc, central, fr, f = S2.setup(L, 5 + L)
    gamma = np.zeros(L * L)
    wf, wg, wv = S2.oracle_steps(pv, pg, pc, central, fr, f, gamma, nsteps, seed=77, trial=3, step0=10)
    ops = tm.ConsumerOps(L)
    ops.step2_setup(*S2.product_params(L, par), central, fr)
    ops.step2_set_strict_means(True)
    ops.step2_upload(f, gamma)
    ops.step2_run(77, trial=3, step0=10, nsteps=nsteps)
    gf, gg, gv, means, ovf = ops.step2_download()
    assert ovf == 0
    assert np.array_equal(gf, wf.reshape(gf.shape)), "fields differ in %d cells" % np.count_nonzero(gf != wf.reshape(gf.shape))
    assert np.array_equal(gg, wg) and np.array_equal(gv, wv)
    ms, launches = ops.last_ms()
    assert launches == 6 * nsteps
    ops.step2_set_strict_means(False)      # back to the tree: same state within the documented few ulp of the means
    ops.step2_upload(f, gamma)
    ops.step2_run(77, trial=3, step0=10, nsteps=1)
    tree = ops.step2_download()[3]
    assert abs(tree[0] - O.mean(f[0])) <= 1e-12 * O.mean(f[0])
    ops.close()


@pytest.mark.gpu
@pytest.mark.parametrize("L", [18, 96])
def test_gpu_step3_strict_means_free_running(L):
    par, pv, pg1, pg2, pc, central, fr1, fr2, f = S3.setup(L, 7)
    ops = S3.product_setup(L, par, central, fr1, fr2)
    ops.set_strict_means(True)
    fo, go = f.copy(), np.zeros((2, L * L))
    ops.upload(f, None)
    for s in range(3):
        ops.run(99, 0, s, 1)
        fp, gp, vp, dyn, ovf = ops.download()
        assert ovf == 0 and [dyn[k] for k in range(3)] == [O.mean(fo[k]) for k in range(3)]
        vo = O.step_3sp(pv, pg1, pg2, pc, central, fr1, fr2, O.make_rng(O.SRC_PHILOX, seed=99, trial=0, step=s), fo, go)
        assert np.array_equal(fp, fo) and np.array_equal(gp, go) and np.array_equal(vp, vo), "step %d" % s
    ops.close()
