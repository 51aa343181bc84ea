"""torchrun worker: Rietkerk lattice cut into row slabs over the ranks (one GPU each), halos exchanged
per step with NCCL send/recv; rank 0 checks the assembled result against a single-GPU run bit for bit.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/run_slab_nccl.py [L] [nsteps]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O                      # noqa: E402  (synthetic frame rule only)
import trophic_maelstorm_b200 as tm         # noqa: E402
from trophic_maelstorm_b200 import ensemble as E  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=13, fringe=(L <= 1024))
    row0, nrows = E.slab_rows(L, world, rank)
    lat = tm.Lattice(p, 1, device=local, row0=row0, nrows=nrows)
    lat.seed(99)
    lat.upload(np.ascontiguousarray(frame[:, row0:row0 + nrows][None]))
    ring = E.cuda_slab_ring(lat, dist)
    ring.step(2)                             # warm-up (NCCL channels)
    lat.upload(np.ascontiguousarray(frame[:, row0:row0 + nrows][None]))
    lat.set_step_counter(0)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    ring.step(nsteps)
    lat.sync(); torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    mine = lat.download()[0]
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    if rank == 0:
        got = np.concatenate(parts, axis=1)
        whole = tm.Lattice(p, 1, device=local)
        whole.seed(99)
        whole.upload(frame[None])
        whole.step(nsteps)
        want = whole.download()[0]
        ok = bool(np.array_equal(got, want))
        print("SLAB_NCCL L=%d world=%d steps=%d bit_exact=%s  %.3f ms/step  %.3g site-updates/s" % (
            L, world, nsteps, ok, 1e3 * dt / nsteps, L * L * nsteps / dt), flush=True)
        if not ok:
            sys.exit(1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
