"""The N > 1 path on CPU: world_size-2 gloo process groups run the same slab decomposition and halo
exchange code that bench.py runs over NCCL (veros_b200/decomp.py), with the CPU oracle standing in for
the kernels.  Checks (a) the exchange fills the ghost columns exactly like the reference's
enforce_boundaries / exchange_overlap would (veros/core/utilities.py:8-21, veros/distributed.py:218-326)
and (b) slabs + exchange reproduce the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _bounds(nxg, world, even):
    """Equal-width slabs (the reference's decomposition) or the unequal widths a cost-balanced cut produces."""
    from veros_b200 import decomp

    if even:
        return [decomp.slab_bounds(nxg, world, r) for r in range(world)]
    return decomp.balanced_slab_bounds(np.r_[np.ones(nxg // 2) * 3.0, np.ones(nxg - nxg // 2)], world, min_width=4)


def _worker(rank, world, port, cyclic, out, even=True):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle
    from veros_b200 import decomp, synthetic

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nxg, ny = 16, 12
        oracle.set_num_threads(1)
        x0, x1 = _bounds(nxg, world, even)[rank]
        st = synthetic.make_workload("global_4deg", nx=x1 - x0, ny=ny, x_offset=x0, nx_global=nxg)
        taup1 = int(st["taup1"])
        oracle.isoneutral_step(st)
        temp, salt = torch.from_numpy(st["temp"]), torch.from_numpy(st["salt"])
        k33 = torch.from_numpy(st["K_33"])
        decomp.exchange_halos_x([temp, salt], cyclic=cyclic, level=taup1)  # strided time level
        decomp.exchange_halos_x([k33], cyclic=cyclic)                      # contiguous 3-D field
        out[rank] = dict(temp=temp.numpy().copy(), salt=salt.numpy().copy(), K_33=k33.numpy().copy(), x0=x0, x1=x1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("even", [True, False], ids=["even", "uneven"])
@pytest.mark.parametrize("cyclic", [True, False])
def test_two_slabs_with_halo_exchange_match_single_process(cyclic, even):
    from oracle import oracle
    from veros_b200 import synthetic

    world, nxg, ny = 2, 16, 12
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), cyclic, out, even), nprocs=world, join=True)
    assert even or out[0]["x1"] - out[0]["x0"] != out[1]["x1"] - out[1]["x0"]  # the uneven case really is uneven

    full = synthetic.make_workload("global_4deg", nx=nxg, ny=ny)
    before = {k: full[k].copy() for k in ("temp", "salt", "K_33")}
    taup1 = int(full["taup1"])
    oracle.isoneutral_step(full)
    ref = {k: full[k].copy() for k in ("temp", "salt", "K_33")}
    # what enforce_boundaries does on one process (utilities.py:13-16)
    if cyclic:
        for k in ("temp", "salt"):
            ref[k][-2:, ..., taup1] = ref[k][2:4, ..., taup1]
            ref[k][:2, ..., taup1] = ref[k][-4:-2, ..., taup1]
        ref["K_33"][-2:] = ref["K_33"][2:4]
        ref["K_33"][:2] = ref["K_33"][-4:-2]
    for rank in range(world):
        got = out[rank]
        x0, x1 = got["x0"], got["x1"]
        sl = slice(x0, x1 + 4)
        for k in ("temp", "salt", "K_33"):
            g, r = got[k], ref[k][sl]
            if not cyclic:
                # the outer ghost columns of the end slabs have no neighbour: they keep their values
                lo = 2 if rank == 0 else 0
                hi = g.shape[0] - 2 if rank == world - 1 else g.shape[0]
                g, r = g[lo:hi], r[lo:hi]
            if k == "K_33":
                assert np.array_equal(g, r), (k, rank)
            else:
                assert np.array_equal(g[..., taup1], r[..., taup1]), (k, rank)
                # the other time levels are not exchanged
                assert np.array_equal(got[k][..., (taup1 + 1) % 3], before[k][sl][..., (taup1 + 1) % 3])
