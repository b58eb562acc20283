#!/usr/bin/env python
"""Multi-GPU check of the peer-memory halo exchange against the NCCL one (run under torchrun, one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_peer_halo.py

Every rank fills tracer-shaped fields with rank-dependent values, exchanges one copy with
decomp.exchange_halos_x (NCCL send/recv) and one with decomp.PeerHaloExchange (stores over NVLink), several
times with the edges modified in between, on a ring and on a closed domain, and compares bit for bit.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from veros_b200 import decomp  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for cyclic in (True, False):
        for shape, level in (((12 + rank, 9, 7, 3), 2), ((10, 6, 5), None)):
            g = torch.Generator(device=dev)
            g.manual_seed(100 + rank)
            a = torch.randn(shape, dtype=torch.float64, device=dev, generator=g)
            b = torch.randn(shape, dtype=torch.float64, device=dev, generator=g)
            ra, rb = a.clone(), b.clone()
            ex = decomp.PeerHaloExchange([a, b], level=level, cyclic=cyclic)
            for rep in range(4):
                decomp.exchange_halos_x([ra, rb], cyclic=cyclic, level=level)
                ex()
                torch.cuda.synchronize()
                same = torch.equal(a, ra) and torch.equal(b, rb)
                ok = ok and same
                if not same:
                    print(f"rank {rank}: MISMATCH cyclic={cyclic} shape={shape} rep={rep}", flush=True)
                a[2:-2] += 0.5 * (rank + 1)
                ra[2:-2] += 0.5 * (rank + 1)
            dist.barrier()
            del ex
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("peer halo exchange == NCCL halo exchange on", world, "ranks:", bool(t.item()), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() else 1)


if __name__ == "__main__":
    main()
