"""Multi-GPU plumbing (one process per GPU, torch.distributed): how the path shards.

* Trial ensembles / a-sweeps (BASELINE configs 2-4) are independent units: `shard_trials` deals the
  (a_i, replicate_j) work list round-robin to ranks, each rank advances its trials in one context,
  and `gather_series` brings the per-trial time series (a few kB) to rank 0 in GLOBAL trial order so
  the reference's order-dependent running variance (P/:173, R/:1293) can be applied there.  There is
  no data-path collective.
* One huge lattice (config 5) is cut into row slabs; `SlabRing` runs the per-step halo exchange
  (4 rows x fields per neighbour, NCCL send/recv through torch.distributed) around `tm_step`.
"""
import numpy as np


def shard_trials(n_total, world, rank):
    """Global trial ids owned by `rank` (round-robin, like dealing a-values to OpenMP threads)."""
    return list(range(rank, n_total, world))


def sweep_work_list(a_values, replicates):
    """[(a_index, replicate)] in the reference's order: a-values outer (R/:4519), replicates inner (R/:3501)."""
    return [(i, j) for i in range(len(a_values)) for j in range(replicates)]


def gather_series(local_ids, local_series, n_total, dist=None):
    """local_series: float64 [len(local_ids), ...].  Returns on rank 0 the array [n_total, ...] in global
    trial order (None elsewhere).  `dist` = torch.distributed (initialised) or None for a single process."""
    local_series = np.ascontiguousarray(local_series, dtype=np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = np.empty((n_total,) + local_series.shape[1:])
        out[np.asarray(local_ids, dtype=int)] = local_series
        return out
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = (list(local_ids), torch.from_numpy(local_series))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return None
    out = np.empty((n_total,) + local_series.shape[1:])
    for ids, ser in gathered:
        if len(ids):
            out[np.asarray(ids, dtype=int)] = ser.numpy()
    return out


def slab_rows(L, world, rank):
    """Row range [row0, row0+nrows) of `rank`: contiguous, sizes differ by at most one row."""
    base, extra = divmod(L, world)
    nrows = base + (1 if rank < extra else 0)
    row0 = rank * base + min(rank, extra)
    return row0, nrows


class SlabRing:
    """Per-step halo exchange of a row-slab decomposition (periodic ring of ranks).

    `lat` needs: halo_pack(which)->ptr-like, halo_unpack(which, buf), step(n), sync(); `make_buf()` returns a
    fresh receive buffer and `as_tensor(x)` turns a packed halo / receive buffer into a tensor `dist` can send.
    """

    def __init__(self, lat, dist, make_buf, as_tensor, after_comm=None):
        self.lat, self.dist = lat, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.up, self.down = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        self.from_up, self.from_down = make_buf(), make_buf()
        self.as_tensor = as_tensor
        self.after_comm = after_comm or (lambda: None)

    def exchange(self):
        d = self.dist
        top, bot = self.lat.halo_pack(0), self.lat.halo_pack(1)   # my first / last 4 rows
        self.lat.sync()
        ops = [d.P2POp(d.isend, self.as_tensor(top), self.up),     # my top rows are the upper neighbour's ghost-below
               d.P2POp(d.isend, self.as_tensor(bot), self.down),
               d.P2POp(d.irecv, self.as_tensor(self.from_down), self.down),   # its top rows -> my ghost-below
               d.P2POp(d.irecv, self.as_tensor(self.from_up), self.up)]
        for r in d.batch_isend_irecv(ops):
            r.wait()   # NCCL: stream-ordered (returns at once); gloo: blocks until the data is in place
        self.after_comm()
        self.lat.halo_unpack(0, self.from_up)
        self.lat.halo_unpack(1, self.from_down)

    def step(self, nsteps=1):
        for _ in range(nsteps):
            self.exchange()
            self.lat.step(1)


class DevicePtrView:
    """Zero-copy view of a raw device pointer for torch (``torch.as_tensor(view, device='cuda')``)."""

    def __init__(self, ptr, n_doubles):
        self.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def cuda_slab_ring(lat, dist):
    """SlabRing over a real `Lattice` slab context: NCCL send/recv on the packed halo buffers.

    Everything is enqueued on the context's own CUDA stream (wrapped as a torch ExternalStream), so pack ->
    send/recv -> unpack -> step are ordered by the stream and the host never blocks between steps.
    """
    import torch
    n = lat.halo_bytes() // 8
    stream = torch.cuda.ExternalStream(lat.cuda_stream())

    def make_buf():
        return torch.empty(n, dtype=torch.float64, device="cuda")

    def as_tensor(x):
        if isinstance(x, torch.Tensor):
            return x
        return torch.as_tensor(DevicePtrView(x, n), device="cuda")

    class _Adapter:
        def __init__(self, lat):
            self.lat = lat
        def halo_pack(self, which):
            return self.lat.halo_pack(which)
        def halo_unpack(self, which, buf):
            self.lat.halo_unpack(which, buf.data_ptr())
        def step(self, n):
            self.lat.step(n)
        def sync(self):
            pass  # stream order is enough: the collectives below run on the same stream

    ring = SlabRing(_Adapter(lat), dist, make_buf, as_tensor)
    plain_exchange = ring.exchange

    def exchange_on_ctx_stream():
        with torch.cuda.stream(stream):
            plain_exchange()

    ring.exchange = exchange_on_ctx_stream
    return ring
