"""ctypes view of include/tm_slab.h: the row-slab ring of one huge Rietkerk lattice (BASELINE config 5).

``SlabRing``      one process per GPU (torchrun): rank 0's NCCL id travels through ``torch.distributed``.
``LocalSlabRing`` one process, several slabs: NCCL between distinct GPUs, or the copy-engine transport
                  (``transport=PEER_COPY``), under which one GPU may hold several slabs (the one-GPU tests).
No numerics live here.
"""
import ctypes as C
import os

import numpy as np

from . import lattice as _lat

OVERLAP, SERIAL, NOCOMM = 0, 1, 2
NCCL, PEER_COPY = 0, 1
ID_BYTES = 128
_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

SLAB_ABI_SYMBOLS = [
    "tm_slab_unique_id", "tm_slab_partition", "tm_slab_create", "tm_slab_create_local", "tm_slab_set_transport", "tm_slab_destroy", "tm_slab_last_error",
    "tm_slab_ctx", "tm_slab_rows", "tm_slab_run", "tm_slab_run_local", "tm_slab_sync", "tm_slab_last_ms", "tm_slab_comm_bytes_per_step",
]


def lib():
    """libtm_slab.so (links libtm_lattice.so and libnccl.so.2; inside a torch process import torch first so that the NCCL
    already loaded is the one used)."""
    global _lib
    if _lib is not None:
        return _lib
    _lat.lib()
    # libtm_slab.so needs "libnccl.so.2".  A process can hold one library of that soname: load torch's bundled NCCL first when
    # there is one, so that a later `import torch` finds the version it was built against (the system NCCL is older).
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                break
    except Exception:
        pass
    path = os.path.join(_HERE, "libtm_slab.so")
    if not os.path.exists(path):
        raise _lat.TmError("%s is missing: run __graft_entry__.build()" % path)
    L = C.CDLL(path)
    i, vp, vpp = C.c_int, C.c_void_p, C.POINTER(C.c_void_p)
    ip = C.POINTER(C.c_int)
    sig = {
        "tm_slab_unique_id": (i, [vp]),
        "tm_slab_partition": (None, [i, i, i, ip, ip]),
        "tm_slab_create": (i, [C.POINTER(_lat.Params), i, i, i, i, vp, vpp]),
        "tm_slab_create_local": (i, [C.POINTER(_lat.Params), i, i, ip, i, vpp]),
        "tm_slab_set_transport": (i, [vp, i]),
        "tm_slab_destroy": (i, [vp]),
        "tm_slab_last_error": (C.c_char_p, [vp]),
        "tm_slab_ctx": (vp, [vp]),
        "tm_slab_rows": (i, [vp, ip, ip]),
        "tm_slab_run": (i, [vp, i, i]),
        "tm_slab_run_local": (i, [vpp, i, i, i]),
        "tm_slab_sync": (i, [vp]),
        "tm_slab_last_ms": (i, [vp, C.POINTER(C.c_float)]),
        "tm_slab_comm_bytes_per_step": (i, [vp, C.POINTER(C.c_uint64)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def partition(L, n_ranks, rank):
    r0, n = C.c_int(), C.c_int()
    lib().tm_slab_partition(L, n_ranks, rank, C.byref(r0), C.byref(n))
    return r0.value, n.value


class _Slab:
    """one tm_slab*: rows [row0, row0 + nrows) of the lattice on one GPU"""

    def __init__(self, handle, params, n_trials):
        self._s = C.c_void_p(handle)
        self._l = lib()
        r0, n = C.c_int(), C.c_int()
        self._l.tm_slab_rows(self._s, C.byref(r0), C.byref(n))
        self.row0, self.nrows, self.L, self.n_trials = r0.value, n.value, params.L, n_trials

    def _ck(self, rc):
        if rc != 0:
            raise _lat.TmError("tm_slab call failed (%d): %s" % (rc, self._l.tm_slab_last_error(self._s).decode()))

    def _ctx(self):
        return C.c_void_p(self._l.tm_slab_ctx(self._s))

    def _lck(self, rc):
        if rc != 0:
            raise _lat.TmError("tm_lattice call failed (%d): %s" % (rc, _lat.lib().tm_last_error(self._ctx()).decode()))

    def upload(self, rows):
        """rows: [n_trials][3][nrows][L] (this slab's rows)"""
        a = np.ascontiguousarray(rows, dtype=np.float64)
        assert a.shape == (self.n_trials, 3, self.nrows, self.L), a.shape
        self._lck(_lat.lib().tm_upload(self._ctx(), _lat._dptr(a)))
        self._lck(_lat.lib().tm_sync(self._ctx()))

    def download(self):
        out = np.empty((self.n_trials, 3, self.nrows, self.L))
        self._lck(_lat.lib().tm_download(self._ctx(), _lat._dptr(out)))
        return out

    def seed(self, seed, trial0=0):
        self._lck(_lat.lib().tm_seed(self._ctx(), seed, trial0))

    def set_strict(self, strict=True):
        self._lck(_lat.lib().tm_set_strict(self._ctx(), 1 if strict else 0))

    def reserve_sms(self, n):
        self._lck(_lat.lib().tm_slab_reserve_sms(self._ctx(), n))

    def measure(self):
        st = (_lat.Stats * self.n_trials)()
        self._lck(_lat.lib().tm_measure(self._ctx(), st))
        return st

    def sync(self):
        self._ck(self._l.tm_slab_sync(self._s))

    def last_ms(self):
        ms = C.c_float()
        self._ck(self._l.tm_slab_last_ms(self._s, C.byref(ms)))
        return ms.value

    def comm_bytes_per_step(self):
        n = C.c_uint64()
        self._ck(self._l.tm_slab_comm_bytes_per_step(self._s, C.byref(n)))
        return n.value

    def close(self):
        if self._s:
            self._l.tm_slab_destroy(self._s)
            self._s = C.c_void_p()


class SlabRing(_Slab):
    """This rank's slab of a ring with one process per GPU.  `dist` = initialised torch.distributed (any backend: it only
    carries the 128-byte NCCL id), or None for a single rank."""

    def __init__(self, params, n_trials, device, dist=None):
        L = lib()
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
        ident = (C.c_ubyte * ID_BYTES)()
        if world > 1:
            import torch
            if rank == 0:
                rc = L.tm_slab_unique_id(ident)
                if rc != 0:
                    raise _lat.TmError("tm_slab_unique_id failed: %s" % L.tm_slab_last_error(None).decode())
            t = torch.tensor(list(bytes(ident)), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda(device)
            dist.broadcast(t, 0)
            ident = (C.c_ubyte * ID_BYTES)(*t.cpu().tolist())
        h = C.c_void_p()
        rc = L.tm_slab_create(C.byref(params), n_trials, device, rank, world, ident, C.byref(h))
        if rc != 0:
            raise _lat.TmError("tm_slab_create failed (%d): %s" % (rc, L.tm_slab_last_error(None).decode()))
        super().__init__(h.value, params, n_trials)
        self.rank, self.world = rank, world

    def run(self, nsteps, mode=OVERLAP):
        self._ck(self._l.tm_slab_run(self._s, nsteps, mode))

    def set_transport(self, transport):
        self._ck(self._l.tm_slab_set_transport(self._s, transport))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocalSlabRing:
    """All slabs of a ring in this process: devices[i] holds rows partition(L, n, i)."""

    def __init__(self, params, n_trials, devices, transport=NCCL):
        L = lib()
        n = len(devices)
        hs = (C.c_void_p * n)()
        rc = L.tm_slab_create_local(C.byref(params), n_trials, n, (C.c_int * n)(*devices), transport, hs)
        if rc != 0:
            raise _lat.TmError("tm_slab_create_local failed (%d): %s" % (rc, L.tm_slab_last_error(None).decode()))
        self._hs = hs
        self.slabs = [_Slab(hs[i], params, n_trials) for i in range(n)]
        self.L, self.n_trials = params.L, n_trials

    def upload(self, fields):
        """fields: [n_trials][3][L][L], cut into the slabs' rows"""
        f = np.asarray(fields, dtype=np.float64).reshape(self.n_trials, 3, self.L, self.L)
        for s in self.slabs:
            s.upload(f[:, :, s.row0:s.row0 + s.nrows, :])

    def download(self):
        return np.concatenate([s.download() for s in self.slabs], axis=2)

    def seed(self, seed):
        for s in self.slabs:
            s.seed(seed)

    def set_strict(self, strict=True):
        for s in self.slabs:
            s.set_strict(strict)

    def run(self, nsteps, mode=OVERLAP):
        rc = lib().tm_slab_run_local(self._hs, len(self.slabs), nsteps, mode)
        if rc != 0:
            raise _lat.TmError("tm_slab_run_local failed (%d): %s" % (rc, "; ".join(lib().tm_slab_last_error(s._s).decode() for s in self.slabs)))

    def sync(self):
        for s in self.slabs:
            s.sync()

    def close(self):
        for s in self.slabs:
            s.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
