"""torchrun worker: a Rietkerk lattice cut into row slabs over the ranks (one GPU each); the ring is libtm_slab.so's
(ncclSend / ncclRecv inside the C layer, halo traffic hidden behind the inner rows); torch.distributed only carries the NCCL id and
gathers the result.  Rank 0 checks the assembled lattice against a single-GPU run bit for bit, in every schedule.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/run_slab_nccl.py [L] [nsteps]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O                      # noqa: E402  (synthetic frame rule only)
import trophic_maelstorm_b200 as tm         # noqa: E402
from trophic_maelstorm_b200 import slab as S  # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = tm.rietkerk_params(L)
    frame = O.synthetic_rietkerk_frame(L, seed=13, fringe=(L <= 1024))
    ring = S.SlabRing(p, 1, local, dist)
    mine = np.ascontiguousarray(frame[:, ring.row0:ring.row0 + ring.nrows][None])
    want = None
    if rank == 0:
        whole = tm.Lattice(p, 1, device=local)
        whole.seed(99)
        whole.upload(frame[None])
        whole.step(nsteps)
        want = whole.download()[0]
        whole.close()
    ok = True
    for transport, mode, name in ((S.NCCL, S.OVERLAP, "nccl/overlap"), (S.NCCL, S.SERIAL, "nccl/serial"),
                                  (S.PEER_COPY, S.OVERLAP, "ipc-copy/overlap"), (S.PEER_COPY, S.SERIAL, "ipc-copy/serial")):
        ring.set_transport(transport)
        ring.seed(99)
        tm.lib().tm_set_step_counter(ring._ctx(), 0)
        ring.upload(mine)
        dist.barrier()
        ring.run(nsteps, mode)
        ring.sync()
        ms = ring.last_ms()
        got_mine = ring.download()[0]
        parts = [None] * world if rank == 0 else None
        dist.gather_object(got_mine, parts, dst=0)
        if rank == 0:
            got = np.concatenate(parts, axis=1)
            same = bool(np.array_equal(got, want))
            ok = ok and same
            print("SLAB_NCCL L=%d world=%d steps=%d schedule=%s bit_exact=%s  %.3f ms/step (first run, channels cold)" % (
                L, world, nsteps, name, same, ms / nsteps), flush=True)
    dist.barrier()
    ring.close()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
