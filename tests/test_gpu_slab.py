"""BASELINE config 5 (row slabs + halo ring) through include/tm_slab.h: every schedule and transport reproduces tm_step on the
undivided lattice bit for bit, because a cell's Philox counter is its global index and the band / inner row views are the same
kernels on shifted pointers."""
import numpy as np
import pytest

import oracle_lib as O
import trophic_maelstorm_b200 as tm
from trophic_maelstorm_b200 import slab as S

pytestmark = pytest.mark.gpu


def _whole(p, frames, nsteps, strict, seed=77):
    lat = tm.Lattice(p, frames.shape[0], device=0)
    lat.set_strict(strict)
    lat.seed(seed)
    lat.upload(frames)
    lat.step(nsteps)
    out = lat.download()
    assert lat.sampler_overflow() == 0
    lat.close()
    return out


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("n_slabs,L", [(2, 96), (3, 100), (4, 128)])
def test_local_ring_on_one_gpu_matches_the_undivided_lattice(n_slabs, L, strict):
    """copy-engine transport, all slabs on GPU 0 (uneven slabs at L=100): overlapped and serial schedules, 6 steps, 2 trials"""
    p = tm.rietkerk_params(L)
    frames = np.stack([O.synthetic_rietkerk_frame(L, seed=5 + t) for t in range(2)])
    want = _whole(p, frames, 6, strict)
    for mode in (S.OVERLAP, S.SERIAL):
        ring = S.LocalSlabRing(p, 2, [0] * n_slabs, transport=S.PEER_COPY)
        ring.set_strict(strict)
        ring.seed(77)
        ring.upload(frames)
        ring.run(2, mode)
        ring.run(4, mode)          # a second call continues from the ghosts the first left
        ring.sync()
        got = ring.download()
        ring.close()
        assert np.array_equal(got, want), (mode, np.abs(got - want).max())


def test_overlapped_steps_mix_with_measure_and_plain_calls():
    """tm_measure between overlapped runs (the join orders the main stream after the band stream), schedules alternating"""
    L = 96
    p = tm.rietkerk_params(L)
    frames = O.synthetic_rietkerk_frame(L, seed=3)[None]
    lat = tm.Lattice(p, 1, device=0)
    lat.seed(77); lat.upload(frames)
    ring = S.LocalSlabRing(p, 1, [0, 0], transport=S.PEER_COPY)
    ring.seed(77); ring.upload(frames)
    for k, mode in enumerate((S.OVERLAP, S.SERIAL, S.OVERLAP)):
        lat.step(3)
        ring.run(3, mode)
        st = [s.measure()[0] for s in ring.slabs]
        ref = lat.measure()[0]
        assert sum(s.active[0] for s in st) == ref.active[0]
        w = [s.nrows / L for s in ring.slabs]
        assert abs(sum(wi * s.mean[0] for wi, s in zip(w, st)) - ref.mean[0]) <= 1e-13 * abs(ref.mean[0])
    assert np.array_equal(ring.download(), lat.download())
    ring.close(); lat.close()


def test_slab_too_thin_is_refused():
    p = tm.rietkerk_params(64)
    ring = S.LocalSlabRing(p, 1, [0, 0, 0, 0], transport=S.PEER_COPY)   # 16 rows per slab < 2 bands + 2 halos
    ring.seed(1); ring.upload(O.synthetic_rietkerk_frame(64, seed=3)[None])
    with pytest.raises(tm.TmError):
        ring.run(1, S.OVERLAP)
    ring.close()


def test_local_ring_over_nccl_between_gpus():
    """one process, every visible GPU, NCCL transport (ncclCommInitAll)"""
    n = tm.lib().tm_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    L = 64 * n
    p = tm.rietkerk_params(L)
    frames = O.synthetic_rietkerk_frame(L, seed=9)[None]
    want = _whole(p, frames, 5, False)
    for transport in (S.NCCL, S.PEER_COPY):
        ring = S.LocalSlabRing(p, 1, list(range(n)), transport=transport)
        ring.seed(77); ring.upload(frames)
        ring.run(5, S.OVERLAP)
        ring.sync()
        got = ring.download()
        ring.close()
        assert np.array_equal(got, want), transport
