"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): two ranks, one x-slab each, run the isoneutral
step through OverlappedStepper (boundary strips -> halo exchange || interior) for three model steps with the
time levels rotating in between, once with the peer-memory exchange (veros_b200_halo_put over NVLink) and once
with pack + NCCL send/recv + unpack.  Every output of every rank must equal the single-GPU run of the global
state bit for bit -- including the ghost planes the exchange fills, which is what
veros/core/thermodynamics.py:293-298 (enforce_boundaries -> veros/distributed.py:218-326 exchange_overlap) does
in the reference.  Replaces the manual scripts/check_peer_halo.py of round 1.

Run by hand on a 2-GPU box:  gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu -q
"""
import os
import socket
import tempfile

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

OUT = ("temp", "salt", "dtemp_iso", "dsalt_iso", "P_diss_iso", "Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33")
STEPS = 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def global_state(kind):
    from veros_b200 import synthetic

    if kind == "analytic":  # global_4deg-shaped, cyclic, TEOS-10
        return synthetic.make_workload("global_4deg", nx=32, ny=20)
    # the random stress state (irregular kbot, islands, junk in never-written elements), cyclic in x
    return synthetic.make_workload("bench_1M", nx=24, ny=14, nz=21, eq_of_state_type=3, enable_cyclic_x=True)


def slab_of(st, x0, nxl):
    """Local arrays of the x-slab [x0, x0 + nxl): planes [x0, x0 + nxl + 4) of every array whose first axis is x."""
    N = st["nx"] + 4
    out = {}
    for k, v in st.items():
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == N and k not in ("dyt", "dyu", "cost", "cosu", "dzt", "dzw", "zt"):
            out[k] = np.ascontiguousarray(v[x0:x0 + nxl + 4])
        else:
            out[k] = v
    out["nx"] = nxl
    return out


def bounds_of(kind, even):
    """Interior x-ranges of the two slabs: equal widths, or the unequal widths a cost-balanced cut produces."""
    nx = global_state(kind)["nx"] if even else None
    if even:
        return [(0, nx // 2), (nx // 2, nx)]
    return [(0, 12), (12, 32)] if kind == "analytic" else [(0, 9), (9, 24)]


def _worker(rank, world, port, kind, halo, outdir, even=True):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    from veros_b200 import decomp
    from veros_b200.state import IsoState

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        st = global_state(kind)
        x0, x1 = bounds_of(kind, even)[rank]
        gs = IsoState.from_numpy(slab_of(st, x0, x1 - x0), dev)
        stepper = decomp.OverlappedStepper(gs, cyclic=True, halo=halo)
        for _ in range(STEPS):
            stepper.step()
            torch.cuda.synchronize()
            gs.advance_time()
        dist.barrier()
        np.savez(os.path.join(outdir, f"{kind}_{halo}_{rank}.npz"), x0=x0, x1=x1, **gs.to_numpy(list(OUT)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("even", [True, False], ids=["even", "uneven"])
@pytest.mark.parametrize("halo", ["peer", "nccl"])
@pytest.mark.parametrize("kind", ["analytic", "random"])
def test_two_rank_overlapped_steps_match_single_gpu_bit_for_bit(kind, halo, even):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp

    from veros_b200 import decomp, isoneutral
    from veros_b200.state import IsoState

    world = 2
    with tempfile.TemporaryDirectory() as outdir:
        mp.spawn(_worker, args=(world, _free_port(), kind, halo, outdir, even), nprocs=world, join=True)
        # single-GPU reference: the global state, plain fused step + the cyclic wrap of enforce_boundaries
        st = global_state(kind)
        g = IsoState.from_numpy(st, "cuda:0")
        for _ in range(STEPS):
            lvl = g.variables.taup1_host
            isoneutral.isoneutral_step(g)
            decomp.exchange_halos_x([g.variables.temp, g.variables.salt], cyclic=True, level=lvl)
            torch.cuda.synchronize()
            g.advance_time()
        ref = g.to_numpy(list(OUT))
        for rank in range(world):
            z = np.load(os.path.join(outdir, f"{kind}_{halo}_{rank}.npz"))
            x0, x1 = int(z["x0"]), int(z["x1"])
            for k in OUT:
                got, want = z[k], ref[k][x0:x1 + 4]
                if k in ("temp", "salt"):  # whole local array incl. the exchanged ghost planes, all time levels
                    assert np.array_equal(got, want), (k, rank)
                else:  # ghost planes of the other outputs belong to the neighbour / are never exchanged
                    assert np.array_equal(got[2:-2], want[2:-2]), (k, rank)
