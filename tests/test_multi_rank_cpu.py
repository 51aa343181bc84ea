"""The N>1 host logic on CPU: world_size-2 (and 3) gloo process groups.

* trial sharding + gather in global replicate order (ensembles / a-sweeps: no data-path collective)
* the per-step halo-exchange choreography of the row-slab decomposition (SlabRing), driven with a
  numpy stand-in for the lattice whose update has the same radius-4 dependency cone as the fused
  RK4 step: slabs + exchange must reproduce the whole-lattice update exactly.
"""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def cone4_update(full):
    """periodic 5-point smoothing applied four times: radius-4 dependency, like one fused RK4 step"""
    x = full
    for _ in range(4):
        x = 0.2 * (x + np.roll(x, 1, -2) + np.roll(x, -1, -2) + np.roll(x, 1, -1) + np.roll(x, -1, -1)) + 0.01 * x * x
    return x


class NumpySlab:
    """numpy stand-in with the Lattice slab interface: local rows + 4 ghost rows above/below."""

    def __init__(self, fields, row0, nrows):
        self.n_fields, self.L = fields.shape[0], fields.shape[2]
        self.nrows = nrows
        self.buf = np.zeros((self.n_fields, nrows + 8, self.L))
        self.buf[:, 4:4 + nrows] = fields[:, row0:row0 + nrows]

    def halo_pack(self, which):
        rows = self.buf[:, 4:8] if which == 0 else self.buf[:, self.nrows:self.nrows + 4]
        return torch.from_numpy(np.ascontiguousarray(rows).reshape(-1).copy())

    def halo_unpack(self, which, t):
        rows = t.numpy().reshape(self.n_fields, 4, self.L)
        if which == 0:
            self.buf[:, 0:4] = rows
        else:
            self.buf[:, self.nrows + 4:self.nrows + 8] = rows

    def step(self, n):
        ext = self.buf
        x = ext
        for _ in range(4):   # non-periodic in rows inside the extended slab: only the interior rows are valid afterwards
            up = np.concatenate([x[:, :1], x[:, :-1]], axis=1); dn = np.concatenate([x[:, 1:], x[:, -1:]], axis=1)
            x = 0.2 * (x + up + dn + np.roll(x, 1, -1) + np.roll(x, -1, -1)) + 0.01 * x * x
        self.buf[:, 4:4 + self.nrows] = x[:, 4:4 + self.nrows]

    def sync(self):
        pass

    def interior(self):
        return self.buf[:, 4:4 + self.nrows].copy()


def _worker(rank, world, port, L, nsteps, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from trophic_maelstorm_b200 import ensemble as E
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # ---- ensembles: shard, compute a per-trial series, gather in global order -------------------------
    n_total = 7
    ids = E.shard_trials(n_total, world, rank)
    series = np.array([[100.0 * t + k for k in range(5)] for t in ids]).reshape(len(ids), 5)
    full = E.gather_series(ids, series, n_total, dist)
    if rank == 0:
        assert np.array_equal(full, np.array([[100.0 * t + k for k in range(5)] for t in range(n_total)]))
    # ---- slabs: halo exchange choreography -----------------------------------------------------------------
    fields = np.random.RandomState(3).random_sample((2, L, L))
    row0, nrows = E.slab_rows(L, world, rank)
    slab = NumpySlab(fields, row0, nrows)
    n = 2 * 4 * L
    ring = E.SlabRing(slab, dist, make_buf=lambda: torch.empty(n, dtype=torch.float64), as_tensor=lambda x: x)
    ring.step(nsteps)
    np.save(os.path.join(out_dir, "slab_%d.npy" % rank), slab.interior())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_gather_and_halo_ring(tmp_path, world):
    L, nsteps = 24, 3
    mp.spawn(_worker, args=(world, _free_port(), L, nsteps, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / ("slab_%d.npy" % r)) for r in range(world)], axis=1)
    want = np.random.RandomState(3).random_sample((2, L, L))
    for _ in range(nsteps):
        want = cone4_update(want)
    assert np.allclose(got, want, rtol=0, atol=1e-15) and np.array_equal(got, want)


def test_shard_helpers():
    sys.path.insert(0, ROOT)
    from trophic_maelstorm_b200 import ensemble as E
    assert E.shard_trials(7, 3, 1) == [1, 4]
    assert E.sweep_work_list([0.1, 0.2], 2) == [(0, 0), (0, 1), (1, 0), (1, 1)]
    rows = [E.slab_rows(10, 4, r) for r in range(4)]
    assert rows == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sum(n for _, n in rows) == 10
