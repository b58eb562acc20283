#!/usr/bin/env python
"""Benchmark of the isoneutral step (pre + isoneutral_diffusion(temp) + (salt), column solves included).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

One "step" = one pass of the hot path (veros/core/thermodynamics.py:430-432) over one synthetic state; unit of
work = one interior grid cell through one step ("cell-update").  Prints ONE JSON line (rank 0).

Defaults follow BASELINE.json's north star:
  * N = 1: the 1 degree global grid (global_1deg, 360 x 160 x 115, TEOS-10, energy on) -- the grid the roofline
    target is stated on.  `also` carries the other single-GPU configs (bench_1M, global_4deg through a CUDA
    graph, the ACC grid, the stand-alone column solve on the shapes of benchmarks/tdma_benchmark.py).
  * N > 1 (torch.distributed.run, one rank per GPU): STRONG scaling of the 0.25 degree grid (global_025deg,
    1440 x 720 x 80 split into N x-slabs of one consistent global state), the 2-cell tracer halo exchange by
    peer-memory stores over NVLink, overlapped with the interior compute (boundary strips first), plus a
    bit-for-bit check of the seam between ranks 0 and 1 against a single-GPU run of the same planes.
  * --impl reference: the UNMODIFIED reference (NumPy backend) from baseline/_ref on the host cores, on a
    bounded x-slab sample of the same workload (baseline/numpy_reference.py); rank 0 only.
See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "isoneutral+TDMA cell-updates/s fp64"
UNIT = "cell-updates/s"
L2_BYTES = 126e6
PRE_BYTES_PER_CELL = 180  # isoneutral_diffusion_pre alone: 28 B read + 152 B written
SOLVE_BYTES_PER_CELL = 42  # stand-alone solve_implicit: a, b, c, d read + out written + two mask bytes


def algorithmic_bytes_per_cell(energy):
    """fused step 244 B/cell (energy off) / 276 B/cell (energy on, reference default)."""
    return 276 if energy else 244


def default_workload(world):
    return "global_1deg" if world == 1 else "global_025deg"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        with open(self.path) as f:
            for line in f:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    smax = float(p[2])
                except ValueError:
                    continue
                for name, val in zip(names, p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


_RESULT_FD = None


def claim_stdout():
    """Keep stdout for the one JSON line: native libraries (NCCL prints its version banner there) and stray
    prints go to stderr for the rest of the run."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------- CPU legs
def sample_kwargs(name, target_cells=1_000_000):
    """A bounded sample of a named workload for the CPU legs: the x-slab [0, nxs) of the same global state
    (all y, all z, ghost planes included), about `target_cells` interior cells."""
    from veros_b200 import synthetic

    cfg = synthetic.WORKLOADS[name]
    per_plane = cfg["ny"] * cfg["nz"]
    nxs = max(8, min(cfg["nx"], int(round(target_cells / per_plane))))
    if name == "bench_1M" or nxs == cfg["nx"]:
        return {}, cfg["nx"]
    return dict(nx=nxs, x_offset=0, nx_global=cfg["nx"]), nxs


def port_baseline(name, kw, seconds=12.0):
    """The CPU oracle (oracle/iso_oracle.c, a C/OpenMP restatement of the reference's NumPy path pinned to its
    golden vectors) on ALL host cores: repeated full steps until `seconds` have passed."""
    from oracle import oracle
    from veros_b200 import synthetic

    oracle.set_num_threads(host_cores())  # torchrun exports OMP_NUM_THREADS=1 to its workers
    cores = oracle.num_threads()
    st = synthetic.make_workload(name, **kw)
    cells = st["nx"] * st["ny"] * st["nz"]
    oracle.isoneutral_step(st)  # warm-up (page faults, thread pool)
    n, t0, times = 0, time.perf_counter(), []
    while True:
        t1 = time.perf_counter()
        oracle.isoneutral_step(st)
        times.append(time.perf_counter() - t1)
        n += 1
        if time.perf_counter() - t0 > seconds or n >= 400:
            break
    return {
        "value": cells / min(times), "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{n} full steps of {name} {st['nx']}x{st['ny']}x{st['nz']} ({cells} cells"
                  f"{', x-slab of the global grid' if kw else ''}), best of {n}; C/OpenMP restatement of the reference "
                  f"NumPy path (oracle/), OpenMP over x-planes",
        "mean_value": cells * n / sum(times),
    }


def numpy_reference_baseline(name, steps, warmup, target_cells):
    """The real reference (NumPy backend, baseline/_ref) on an x-slab sample of the workload; None if the
    reference was not installed into baseline/_ref."""
    from baseline import numpy_reference as nr
    from veros_b200 import synthetic

    if not nr.available():
        return None
    kw, nxs = sample_kwargs(name, target_cells)
    st = synthetic.make_workload(name, **kw)
    cells = st["nx"] * st["ny"] * st["nz"]
    best, mean = nr.time_steps(st, steps, warmup)
    return {
        "value": cells / mean, "unit": UNIT, "cores": 1, "kind": "reference",
        "sample": f"{steps} timed steps (+{warmup} warm-up) of the unmodified reference, NumPy backend (baseline/_ref: "
                  f"isoneutral_diffusion_pre + isoneutral_diffusion(temp) + (salt), SciPy dgtsv column solve), on the x-slab "
                  f"[0, {nxs}) of {name} = {st['nx']}x{st['ny']}x{st['nz']} ({cells} cells); single-threaded by construction "
                  f"(NumPy ufuncs + LAPACK dgtsv); mean of the timed steps; JAX-CPU not measurable (jax is not installed)",
        "best_value": cells / best, "ms_per_step": mean * 1e3, "cells": cells,
    }


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return
    name = args.workload
    base = numpy_reference_baseline(name, args.steps, args.warmup, target_cells=1_000_000)
    port = None
    try:
        pkw, _ = sample_kwargs(name, 8_000_000)
        port = port_baseline(name, pkw, seconds=6.0)
    except Exception as err:  # the port is context here, never the headline
        port = {"error": str(err)}
    if base is None:  # baseline/_ref absent: fall back to the port (kind says so)
        base = port
        ms = None
    else:
        ms = base["ms_per_step"]
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if world == 1 else args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "note": "host cores only; each step is a bounded x-slab sample of the same workload "
                   "the GPU arm runs (throughput per cell does not depend on the slab width)"},
        "cpu_baseline": base, "oracle_port_all_cores": port,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------- helpers (GPU arm)
def bind_to_gpu_numa_node(local_rank):
    """Pin this rank (and therefore its pinned staging buffers, first touch) to the NUMA node of its GPU: with 8
    ranks staging through one socket the e2e leg collapses (round 1: 27 % weak efficiency)."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return None
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return node
    except Exception:
        return None
    return None


def timed(fn, n):
    """Average device time of fn() in ms (CUDA events on the current stream)."""
    import torch

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def assemble_on_device(name, nslabs, dev):
    """IsoState of a whole named grid on `dev`, generated on the host slab by slab (x_offset) so that the host never
    holds more than one slab: local planes [0, nxl + 4) of slab s are global planes [s nxl, s nxl + nxl + 4)."""
    import torch

    from veros_b200 import synthetic
    from veros_b200.state import IsoState

    nxg = synthetic.WORKLOADS[name]["nx"]
    nxl = nxg // nslabs
    first = synthetic.make_workload(name, nx=nxl, x_offset=0, nx_global=nxg)
    N = nxg + 4
    proto = IsoState.from_numpy(first, dev)
    full = {}
    for k, t in vars(proto.variables).items():
        if hasattr(t, "shape") and t.dim() >= 1 and t.shape[0] == nxl + 4 and k not in ("dyt", "dyu", "cost", "cosu", "dzt", "dzw", "zt"):
            full[k] = torch.empty((N,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            full[k][:nxl + 4].copy_(t)
    for s_ in range(1, nslabs):
        part = IsoState.from_numpy(synthetic.make_workload(name, nx=nxl, x_offset=s_ * nxl, nx_global=nxg), dev)
        for k in full:
            full[k][s_ * nxl:s_ * nxl + nxl + 4].copy_(getattr(part.variables, k))
        del part
    for k, t in full.items():
        setattr(proto.variables, k, t)
    proto.settings.nx = nxg
    proto._workspace = None
    proto.validate()
    return proto


def side_measurements(dev, peak):
    """The other single-GPU configs of BASELINE.json, each the fused step on its own synthetic state (not part of
    `value`): bench_1M (configs[1]), global_4deg through a CUDA graph (configs[2], latency path), the ACC grid
    (configs[0]) and the stand-alone column solve on the shapes of benchmarks/tdma_benchmark.py:30-39."""
    import numpy as np
    import torch

    from veros_b200 import _lib, isoneutral, synthetic, utilities
    from veros_b200.state import IsoState

    out = {}
    for name, reps, graph in (("bench_1M", 3, False), ("global_4deg", 4, True), ("acc", 4, True)):
        st = synthetic.make_workload(name)
        cells = st["nx"] * st["ny"] * st["nz"]
        states = [IsoState.from_numpy(st, dev) for _ in range(reps)]
        plans = [isoneutral.StepPlan(s) for s in states]
        for p in plans:
            p()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        plans[0]()
        launches = _lib.launch_count() - l0
        ms = timed(lambda k: plans[k % reps](), 50)
        rec = {"grid": f"{st['nx']}x{st['ny']}x{st['nz']}", "eq_of_state_type": int(st["eq_of_state_type"]),
               "ms_per_step": ms, "value": cells / (ms * 1e-3), "kernel_launches_per_step": int(launches),
               "step_roofline_frac": cells * algorithmic_bytes_per_cell(bool(st["enable_conserve_energy"])) / (ms * 1e-3) / 1e9 / peak}
        if graph:
            for p in plans:
                p.capture()
            msg = timed(lambda k: plans[k % reps](), 200)
            rec.update(ms_per_step_cuda_graph=msg, value_cuda_graph=cells / (msg * 1e-3))
        out[name] = rec
        del states, plans
        torch.cuda.empty_cache()
    # the strong-scaling grid on ONE GPU (the N = 1 point of BASELINE.json's 0.25 degree config): 83 M cells, 25 GB state,
    # assembled on the device from sixteen 90-plane slabs of the same generator the multi-GPU ranks use
    try:
        big = assemble_on_device("global_025deg", 16, dev)
        cells = big.settings.nx * big.settings.ny * big.settings.nz
        plan = isoneutral.StepPlan(big)
        for _ in range(2):
            plan()
        ms = timed(lambda k: plan(), 5)
        out["global_025deg"] = {"grid": f"{big.settings.nx}x{big.settings.ny}x{big.settings.nz}", "eq_of_state_type": 5,
                                "ms_per_step": ms, "value": cells / (ms * 1e-3),
                                "step_roofline_frac": cells * 276 / (ms * 1e-3) / 1e9 / peak,
                                "note": "one state of 25 GB streamed per step (far larger than L2); the single-GPU point of the "
                                        "strong-scaling curve that `--gpus N` measures"}
        del big, plan
        torch.cuda.empty_cache()
    except Exception as err:  # e.g. a GPU without the memory
        out["global_025deg"] = {"error": str(err)}
    # stand-alone solve_implicit: random 70 x 60 x 50 systems as the reference's TDMA benchmark, and a 1 degree sized batch
    rng = np.random.default_rng(17)
    for label, (nx, ny, nz) in (("tdma_benchmark_70x60x50", (70, 60, 50)), ("global_1deg_364x164x115", (364, 164, 115))):
        a, b, c, d = (torch.from_numpy(rng.standard_normal((nx, ny, nz))).to(dev) for _ in range(4))
        kbot = rng.integers(0, nz, size=(nx, ny))
        kk = np.arange(nz)[None, None, :]
        water = torch.from_numpy(((kbot > 0)[..., None] & (kk >= (kbot - 1)[..., None]))).to(dev)
        edge = torch.from_numpy(((kbot > 0)[..., None] & (kk == (kbot - 1)[..., None]))).to(dev)
        for _ in range(3):
            utilities.solve_implicit(a, b, c, d, water, edge)
        ms = timed(lambda k: utilities.solve_implicit(a, b, c, d, water, edge), 30)
        n = nx * ny * nz
        out["solve_implicit_" + label] = {"ms": ms, "cells_per_s": n / (ms * 1e-3), "algorithmic_bytes_per_cell": SOLVE_BYTES_PER_CELL,
                                          "achieved_gbs": n * SOLVE_BYTES_PER_CELL / (ms * 1e-3) / 1e9,
                                          "frac": n * SOLVE_BYTES_PER_CELL / (ms * 1e-3) / 1e9 / peak,
                                          "note": "native z-contiguous layout, dgtsv replay incl. pivoting, output allocation included"}
    return out


def neighbour_ops(state, dev, peak, cells):
    """Device time of the SURVEY.md 8f rank 3 / rank 4 ops on the benchmark grid (random inputs of the right shapes;
    not part of `value`): implicit_vert_friction, advect_tempsalt (+ Adams-Bashforth), set_eke_diffusivities,
    isoneutral_diag_streamfunction, each with its algorithmic bytes per cell and fraction of the measured HBM peak."""
    import torch

    from veros_b200 import eke, friction, isoneutral, thermodynamics

    vs, st = state.variables, state.settings
    N, M, nz = st.nx + 4, st.ny + 4, st.nz
    gen = torch.Generator(device=dev)
    gen.manual_seed(11)
    r = lambda *shape, scale=1.0: torch.randn(shape, dtype=torch.float64, device=dev, generator=gen) * scale
    extra = dict(u=r(N, M, nz, 3), v=r(N, M, nz, 3), w=r(N, M, nz, 3, scale=1e-4), kappaM=r(N, M, nz).abs_() * 1e-2,
                 du_mix=r(N, M, nz), dv_mix=r(N, M, nz), K_diss_v=r(N, M, nz), area_t=r(N, M).abs_() + 1.0,
                 area_v=r(N, M).abs_() + 1.0, dtemp=r(N, M, nz, 3, scale=1e-6), dsalt=r(N, M, nz, 3, scale=1e-7),
                 Nsqr=r(N, M, nz, 3, scale=1e-5), eke=r(N, M, nz, 3, scale=1e-2), coriolis_t=r(N, M, scale=1e-4),
                 beta=r(N, M).abs_() * 2e-11, B1_gm=r(N, M, nz), B2_gm=r(N, M, nz))
    saved = {k: getattr(vs, k, None) for k in extra}
    saved_tr = (vs.temp.clone(), vs.salt.clone())
    for k, v_ in extra.items():
        setattr(vs, k, v_)
    st.dt_mom, st.AB_eps, st.enable_superbee_advection = float(st.dt_tracer), 0.1, True
    st.enable_eke, st.enable_eke_isopycnal_diffusion = True, True
    st.pi, st.eke_lmin, st.eke_cross, st.eke_crhin, st.eke_k_max, st.eke_c_k = 3.141592653589793, 100.0, 2.0, 1.0, 1e4, 0.4
    st.K_gm_0, st.K_iso_0 = 1000.0, 1000.0
    kg, ki = vs.K_gm.clone(), vs.K_iso.clone()
    ops = {
        "implicit_vert_friction": (lambda: friction.implicit_vert_friction(state), 90,
                                   "veros/core/friction.py:92-205, coefficient assembly fused into two dgtsv solves (3 launches)"),
        "advect_tempsalt_adams_bashforth": (lambda: thermodynamics.advect_tempsalt(state), 92,
                                            "thermodynamics.py:10-62,223-245 with superbee fluxes, both tracers, one launch"),
        "set_eke_diffusivities": (lambda: eke.set_eke_diffusivities_kernel(state), 57, "veros/core/eke.py:34-85, one launch"),
        "isoneutral_diag_streamfunction": (lambda: isoneutral.isoneutral_diag_streamfunction(state), 88,
                                           "isoneutral.py:232-258, one launch"),
    }
    out = {}
    for name_, (fn, nbytes, note) in ops.items():
        for _ in range(3):
            fn()
        ms = timed(lambda k: fn(), 10)
        gbs = cells * nbytes / (ms * 1e-3) / 1e9
        out[name_] = {"ms": ms, "algorithmic_bytes_per_cell": nbytes, "achieved_gbs": gbs, "frac": gbs / peak, "note": note}
    vs.K_gm.copy_(kg)
    vs.K_iso.copy_(ki)
    vs.temp.copy_(saved_tr[0])
    vs.salt.copy_(saved_tr[1])
    for k, v_ in saved.items():
        if v_ is None:
            delattr(vs, k)
        else:
            setattr(vs, k, v_)
    return out


def seam_parity(args, name, kw, xb, nxg, rank, world, dev, cyclic, halo, overlap):
    """One step of fresh slab states on all ranks (same stepper class as the timed run), then ranks 0 and 1
    compare the planes either side of their common boundary -- the exchanged ghost planes included -- with a
    single-GPU run of those planes (a 16-plane slab around the seam generated with the same x_offset mechanism).
    Returns a dict on rank 0, None elsewhere."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from veros_b200 import decomp, isoneutral, synthetic
    from veros_b200.state import IsoState

    st = synthetic.make_workload(name, **kw)  # this rank's slab; `xb` = first interior plane of rank 1 (global index)
    lvl = int(st["taup1"])
    gs = IsoState.from_numpy(st, dev)
    if overlap:
        decomp.OverlappedStepper(gs, cyclic=cyclic, halo=halo).step()
    else:
        isoneutral.isoneutral_step(gs)
        make = decomp.PeerHaloExchange if halo == "peer" else decomp.TracerHaloExchange
        make([gs.variables.temp, gs.variables.salt], level=lvl, cyclic=cyclic)()
    torch.cuda.synchronize()
    H = 8  # planes compared on each side
    vs = gs.variables
    names = ("temp", "salt", "dtemp_iso", "dsalt_iso", "K_33", "K_11")

    def pick(t, sl):
        t = t[sl]
        return (t[..., lvl] if t.dim() == 4 else t).contiguous()

    result = None
    if rank == 1:
        # west side of rank 1: ghost planes [0, 2) (filled by the exchange) + first H interior planes
        buf = torch.stack([pick(getattr(vs, n), slice(0, 2 + H)) for n in names])
        dist.send(buf, dst=0)
    if rank == 0:
        theirs = torch.empty((len(names), 2 + H) + tuple(vs.K_33.shape[1:]), dtype=torch.float64, device=dev)
        dist.recv(theirs, src=1)
        mine = torch.stack([pick(getattr(vs, n), slice(vs.K_33.shape[0] - 2 - H, vs.K_33.shape[0])) for n in names])
        # single-GPU reference: interior planes [nxl - H, nxl + H) of the global grid as one small slab
        seam = synthetic.make_workload(name, nx=2 * H, x_offset=xb - H, nx_global=nxg)
        ss = IsoState.from_numpy(seam, dev)
        isoneutral.isoneutral_step(ss)
        torch.cuda.synchronize()
        ok, worst = True, {}
        for q, n in enumerate(names):
            ref = pick(getattr(ss.variables, n), slice(None))          # local planes 0 .. 2H+4 (ghosts + 2H interior)
            # rank 0: its last H interior planes = seam interior [2, 2+H); its east ghosts = seam [2+H, 4+H) (tracers only)
            a = torch.equal(mine[q][:H], ref[2:2 + H])
            # rank 1: its first H interior planes = seam [2+H, 2+2H); its west ghosts = seam [H, 2+H) (tracers only)
            b = torch.equal(theirs[q][2:], ref[2 + H:2 + 2 * H])
            g = True
            if n in ("temp", "salt"):
                g = torch.equal(mine[q][H:], ref[2 + H:4 + H]) and torch.equal(theirs[q][:2], ref[H:2 + H])
            ok = ok and a and b and g
            worst[n] = bool(a and b and g)
        result = {"bit_identical": bool(ok), "fields": worst,
                  "checked": f"one step on fresh slabs; {H} interior planes either side of the rank 0 / rank 1 boundary and the "
                             f"exchanged ghost planes of temp/salt[taup1] against a single-GPU step of global planes "
                             f"[{xb - H}, {xb + H})"}
    dist.barrier()
    return result


# ---------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="default: global_1deg on one GPU, global_025deg on several")
    ap.add_argument("--replicas", type=int, default=0, help="state replicas rotated through (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the side measurements under `also`")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (forced above 20 M cells per rank)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="N > 1: halo exchange by stores into the neighbours' memory over NVLink (one kernel per rank, CUDA "
                         "IPC mappings) or by pack + NCCL send/recv + unpack")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = the named grid is split over the ranks (default, BASELINE.json's 0.25 degree "
                         "config); weak = every rank gets a slab of the named size")
    ap.add_argument("--decomp", default="balanced", choices=["balanced", "even"],
                    help="N > 1, strong scaling: x-slab cuts that equalise the cost per rank (default) or equal widths as the "
                         "reference's decomposition (veros/distributed.py:124-128)")
    ap.add_argument("--profile", action="store_true",
                    help="only warm-up + timed steps + per-kernel pass (for ncu launch lists): no e2e, cpu or extra legs")
    ap.add_argument("--overlap", default="auto", choices=["auto", "on", "off"],
                    help="N > 1: hide the halo exchange behind interior compute (auto: only with --halo nccl on slabs of "
                         ">= 3 M cells; the peer-memory exchange is too short to be worth the extra strip passes)")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3)
    if args.profile:
        args.no_cpu = args.no_extra = True

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        args.workload = default_workload(max(world, args.gpus))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from veros_b200 import _lib, decomp, isoneutral, synthetic
    from veros_b200.host import HostStepper
    from veros_b200.state import IsoState

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        import datetime

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a collective that some rank never joins must fail, not hang the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    name, kw = args.workload, {}
    scaling = "weak"
    if world > 1:
        scaling = args.scaling
        if name == "bench_1M":
            scaling = "weak"  # the random stress state has no global version: unrelated slabs, timing only
            kw = dict(seed=17 + rank)
        else:
            nxl = synthetic.WORKLOADS[name]["nx"]
            if scaling == "strong":
                # one consistent global state cut into x-slabs; by default the cuts equalise the per-slab COST (wet cells
                # + 0.34 x dry cells), not the width: continents make equal-width slabs unequal (scaling_breakdown)
                nxg = nxl
                if args.decomp == "balanced":
                    bounds = decomp.balanced_slab_bounds(synthetic.analytic_plane_costs(name), world)
                else:
                    bounds = [decomp.slab_bounds(nxg, world, r) for r in range(world)]
                x0, x1 = bounds[rank]
                kw = dict(nx=x1 - x0, x_offset=x0, nx_global=nxg)
            else:
                nxg = nxl * world
                bounds = [(r * nxl, (r + 1) * nxl) for r in range(world)]
                kw = dict(nx=nxl, x_offset=rank * nxl, nx_global=nxg)
    st = synthetic.make_workload(name, **kw)
    nx, ny, nz = st["nx"], st["ny"], st["nz"]
    cells = nx * ny * nz
    energy = bool(st["enable_conserve_energy"])
    cyclic = bool(st.get("enable_cyclic_x", True))

    # state replicas: consecutive steps touch different memory, so nothing is served from L2
    probe = IsoState.from_numpy(st, dev)
    state_bytes = sum(t.numel() * t.element_size() for t in vars(probe.variables).values() if hasattr(t, 'numel'))
    # EVERY decision that changes which collectives / exchanges a rank takes part in is made on values all ranks
    # agree on (unequal slab widths give the ranks different sizes): the largest slab decides
    cells_max = int(max_over_ranks(float(cells)))
    state_bytes_max = max_over_ranks(float(state_bytes))
    replicas = args.replicas or max(2, min(8, int(3 * L2_BYTES / state_bytes_max) + 1))
    if state_bytes_max > 4e9 and not args.replicas:
        replicas = 1  # one pass over the state already streams several L2 sizes
    states = [probe] + [IsoState.from_numpy(st, dev) for _ in range(replicas - 1)]

    plans = [isoneutral.StepPlan(s) for s in states]  # argument marshalling done once, as under XLA
    # Overlap (boundary strips first, exchange hidden behind the interior) pays when the exchange is slow relative to
    # the step: measured on 2 GPUs, 0.25 degree strong scaling (profiles/r02_scaling.md): peer-memory exchange after
    # the step 7.23 ms/step (98 % efficient) vs overlapped 8.19 ms (the two extra strip passes cost more than the
    # ~20 us exchange they hide); with pack + NCCL + unpack on >= 3 M-cell slabs the overlap wins (round 1).
    overlap = world > 1 and (args.overlap == "on" or (args.overlap == "auto" and args.halo == "nccl" and cells_max >= 3_000_000))

    def build_exchange(halo):
        steppers_ = [decomp.OverlappedStepper(s, cyclic=cyclic, halo=halo) for s in states] if overlap else None
        make_exchange = decomp.PeerHaloExchange if halo == "peer" else decomp.TracerHaloExchange
        exchanges_ = [make_exchange([s.variables.temp, s.variables.salt], level=int(st["taup1"]), cyclic=cyclic)
                      for s in states] if (world > 1 and not overlap) else None
        return steppers_, exchanges_

    halo_note = ""
    try:
        steppers, exchanges = build_exchange(args.halo)
    except decomp.PeerSetupError as err:  # raised on all ranks alike (e.g. allocator blocks that cannot be exported)
        args.halo, halo_note = "nccl", f" (peer-memory mapping unavailable: {err})"
        steppers, exchanges = build_exchange("nccl")

    def step(q):
        if world == 1:
            plans[q]()
        elif steppers is not None:
            steppers[q].step()  # boundary strips -> exchange || interior
        else:
            plans[q]()
            exchanges[q]()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for w in range(args.warmup):
        step(w % replicas)
    barrier()
    # ---- timed region: EXACTLY K steps -------------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    barrier()
    e0.record()
    for k in range(args.steps):
        step(k % replicas)
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    total_cells = cells if world == 1 else (world * cells if name == "bench_1M" else nxg * ny * nz)
    value = total_cells * args.steps / (ms_total * 1e-3)

    # ---- N > 1: where the step time goes -- compute alone on every rank (load balance) and the exchange alone ----
    scaling_diag = None
    if world > 1:
        nrep = min(args.steps, 10)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for k in range(nrep):
            plans[k % replicas]()
        c1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([c0.elapsed_time(c1) / nrep], dtype=torch.float64, device=dev)
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        x_ms = None
        if exchanges is not None:
            barrier()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record()
            for k in range(nrep):
                exchanges[k % replicas]()
            x1.record()
            barrier()
            x_ms = max_over_ranks(x0.elapsed_time(x1) / nrep)
        wet = float(st["maskT"][2:-2, 2:-2].mean())
        wt = torch.tensor([wet], dtype=torch.float64, device=dev)
        allw = [torch.zeros_like(wt) for _ in range(world)]
        dist.all_gather(allw, wt)
        scaling_diag = {"compute_only_ms_per_rank": [round(float(t.item()), 4) for t in allc],
                        "exchange_only_ms": x_ms, "wet_cell_fraction_per_rank": [round(float(t.item()), 3) for t in allw],
                        "note": "step time = slowest rank's compute + exchange + hand-shake skew; equal-width x-slabs of a grid "
                                "with continents are not equally expensive"}

    # ---- per-kernel times INSIDE the fused call: the library records the caller's CUDA events between its
    # kernels (veros_b200_profile_events): [0] start, [1] before / [2] after the slope+flux kernel, [3] end
    import ctypes

    L = _lib.lib()
    ksteps = min(args.steps, 20)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(ksteps)]
    for row in evs:
        for e in row:
            e.record()  # creates the underlying cudaEvent_t
    torch.cuda.synchronize()
    for k in range(ksteps):
        handles = (ctypes.c_void_p * 4)(*[e.cuda_event for e in evs[k]])
        L.veros_b200_profile_events(handles, 4)
        plans[k % replicas]()
        torch.cuda.synchronize()
    L.veros_b200_profile_events(None, 0)
    k_pre = "iso_pre_kernel (slopes + tensor + fluxes)"
    k_upd = "update_kernel (divergence + column solve + tendencies + dissipation)"
    kern_ms = {
        "setup (+eos5)": sum(r[0].elapsed_time(r[1]) for r in evs) / ksteps,
        k_pre: sum(r[1].elapsed_time(r[2]) for r in evs) / ksteps,
        k_upd: sum(r[2].elapsed_time(r[3]) for r in evs) / ksteps,
    }

    # ---- the three stand-alone ops of the reference call surface, for context -----------------------------
    names = ("pre", "diffusion_temp", "diffusion_salt")
    oev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(ksteps)]
    for k in range(ksteps):
        s_ = states[k % replicas]
        vs = s_.variables
        oev[k][0].record()
        isoneutral.isoneutral_diffusion_pre(s_)
        oev[k][1].record()
        isoneutral.isoneutral_diffusion(s_, vs.temp, True)
        oev[k][2].record()
        isoneutral.isoneutral_diffusion(s_, vs.salt, False)
        oev[k][3].record()
    torch.cuda.synchronize()
    op_ms = {n: sum(e[q].elapsed_time(e[q + 1]) for e in oev) / ksteps for q, n in enumerate(names)}

    # ---- the step after the path: vertmix_tempsalt on the same tracers -------------
    from veros_b200 import thermodynamics
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    N_, M_ = nx + 4, ny + 4
    vmix_fields = dict(
        kappaH=torch.rand((N_, M_, nz), dtype=torch.float64, device=dev, generator=gen) * 1e-3,
        forc_temp_surface=(torch.rand((N_, M_), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 1e-5,
        forc_salt_surface=(torch.rand((N_, M_), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 1e-6,
    )
    for s_ in states:
        for k_, v_ in vmix_fields.items():
            setattr(s_.variables, k_, v_)
    cyc_saved = [s_.settings.enable_cyclic_x for s_ in states]
    for s_ in states:
        s_.settings.enable_cyclic_x = False  # kernel only: the exchange is timed with the step above
    vev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(ksteps)]
    if world == 1:
        thermodynamics.vertmix_tempsalt(states[0])  # the reference-facing call once (allocates the tendencies)
    else:
        states[0].variables.dtemp_vmix = torch.empty((N_, M_, nz), dtype=torch.float64, device=dev)
        states[0].variables.dsalt_vmix = torch.empty((N_, M_, nz), dtype=torch.float64, device=dev)
    for s_ in states[1:]:
        s_.variables.dtemp_vmix, s_.variables.dsalt_vmix = states[0].variables.dtemp_vmix, states[0].variables.dsalt_vmix
    vplans = [thermodynamics.VertmixPlan(s_) for s_ in states]
    for k in range(ksteps):
        vev[k][0].record()
        vplans[k % replicas]()
        vev[k][1].record()
    torch.cuda.synchronize()
    for s_, c_ in zip(states, cyc_saved):
        s_.settings.enable_cyclic_x = c_
    vmix_ms = sum(e[0].elapsed_time(e[1]) for e in vev) / ksteps
    VMIX_BYTES = 56  # R temp,salt@taup1 16 + kappaH 8; W temp,salt 16 + dtemp_vmix,dsalt_vmix 16

    peak, peak_src = load_peaks()
    step_bytes = algorithmic_bytes_per_cell(energy)
    step_gbs = cells * step_bytes / (ms_step * 1e-3) / 1e9
    # dominant kernel of the step; algorithmic bytes: isoneutral_diffusion_pre 180 B/cell, the rest of the fused
    # step (276 or 244 B/cell) belongs to the update kernel
    dom, dom_bytes, dom_key = (k_pre, PRE_BYTES_PER_CELL, "iso_pre_kernel") if kern_ms[k_pre] >= kern_ms[k_upd] else \
        (k_upd, step_bytes - PRE_BYTES_PER_CELL, "update_kernel")
    dom_gbs = cells * dom_bytes / (kern_ms[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{name}:{dom_key}")
    roofline = {
        "kernel": dom, "bound": "hbm", "achieved": dom_gbs, "peak": peak,
        "unit": "GB/s", "frac": dom_gbs / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_cell": dom_bytes, "ms_per_launch": kern_ms[dom],
        "note": "CUDA events recorded by the library around its kernels inside the fused step call "
                "(veros_b200_profile_events); the slope kernel runs as two launches (east+north / top faces) on large "
                "grids and is timed as one; traffic = ncu dram bytes per step of that kernel (profiles/traffic.json)",
    }
    step_roofline = {"bytes_per_cell": step_bytes, "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                     "frac": step_gbs / peak, "kernels_ms": kern_ms, "standalone_ops_ms": op_ms}
    next_ops = {"vertmix_tempsalt": {"ms": vmix_ms, "algorithmic_bytes_per_cell": VMIX_BYTES,
                                     "achieved_gbs": cells * VMIX_BYTES / (vmix_ms * 1e-3) / 1e9,
                                     "frac": cells * VMIX_BYTES / (vmix_ms * 1e-3) / 1e9 / peak,
                                     "note": "the step after the path (SURVEY.md 8f rank 1), one kernel, not part of `value`"}}
    if world == 1 and not args.profile:
        next_ops.update(neighbour_ops(states[0], dev, peak, cells))
    if args.profile:
        if rank == 0:
            sampler.stop()
            emit({"profile_only": True, "ms_per_step": ms_step, "kernels_ms": kern_ms})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- multi-GPU: seam parity against a single-GPU run (fresh states, outside the timed region) --------
    del states[1:], plans[1:]
    steppers = exchanges = vplans = None
    torch.cuda.empty_cache()
    parity = None
    if world > 1 and name != "bench_1M":
        parity = seam_parity(args, name, kw, bounds[0][1], nxg, rank, world, dev, cyclic, args.halo, overlap)

    # ---- strong scaling: the SAME grid on one GPU (rank 0 alone, the others wait), so that the line carries its own
    # single-GPU reference point -------------------------------------------------------------------------------------
    one_gpu = None
    if world > 1 and scaling == "strong" and not args.no_extra:
        if rank == 0:
            try:
                del states[:], plans[:]
                torch.cuda.empty_cache()
                big = assemble_on_device(name, 16 if nxg % 16 == 0 else 1, dev)
                plan = isoneutral.StepPlan(big)
                for _ in range(2):
                    plan()
                ms1 = timed(lambda k: plan(), 5)
                one_gpu = {"ms_per_step": ms1, "value": nxg * ny * nz / (ms1 * 1e-3),
                           "note": f"the whole {nxg}x{ny}x{nz} grid on rank 0's GPU alone, same kernels, no exchange"}
                del big, plan
                torch.cuda.empty_cache()
            except Exception as err:
                one_gpu = {"error": str(err)}
        dist.barrier()

    # ---- end to end: host buffers in, host buffers out ----------------------------------------------
    hs = None
    if args.no_e2e or cells_max > 20_000_000:
        clocks = sampler.stop() if rank == 0 else None
        e2e = None  # the pinned staging buffers of this leg would be tens of GB
    else:
        del states[:], plans[:]
        torch.cuda.empty_cache()
        hs = HostStepper(st, dev)
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            hs.step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hs.step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        clocks = sampler.stop() if rank == 0 else None  # sampled from warm-up to here (all under load)
        e2e = {"value": total_cells * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": hs.h2d_bytes,
               "d2h_bytes_per_step": hs.d2h_bytes, "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
               "note": "HostStepper.step(): pinned host -> device copies of the step's inputs, the fused step, device -> host "
                       "copies of all twelve outputs, slab-pipelined over three streams" +
                       (f"; rank pinned to NUMA node {numa} of its GPU" if numa is not None else "")}
        del hs
        torch.cuda.empty_cache()

    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        extra = side_measurements(dev, peak)

    base = ref_np = None
    if rank == 0 and world == 1 and not args.no_cpu:
        base = port_baseline(name, kw)
        try:
            ref_np = numpy_reference_baseline(name, steps=3, warmup=1, target_cells=500_000)
        except Exception as err:
            ref_np = {"error": str(err)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": name, "nx": nx, "ny": ny, "nz": nz, "cells_per_gpu": cells,
                "global_grid": f"{nxg if (world > 1 and name != 'bench_1M') else nx}x{ny}x{nz}",
                "slab_widths": [b - a for a, b in bounds] if (world > 1 and name != "bench_1M") else None,
                "decomposition": (args.decomp + " x-slabs") if (world > 1 and scaling == "strong") else "even x-slabs",
                "eq_of_state_type": int(st["eq_of_state_type"]), "enable_conserve_energy": energy,
                "parallelism": f"x-slabs x{world}" + (((" + ring halo exchange of temp/salt[taup1] by peer-memory stores over NVLink"
                                                        if args.halo == "peer" else " + NCCL ring halo exchange of temp/salt[taup1]") +
                                (", boundary strips first, exchange overlapped with interior" if overlap else ", exchange after the step") + halo_note)
                                if world > 1 else ""),
                "l2": (f"inputs larger than L2: {replicas} state replicas of {state_bytes / 1e6:.0f} MB rotated, "
                       f"no replica is touched twice in a row") if replicas > 1 else
                      f"inputs larger than L2: one state of {state_bytes / 1e6:.0f} MB streamed per step",
            },
            "roofline": roofline, "step_roofline": step_roofline, "next_ops": next_ops, "cpu_baseline": base,
            "numpy_reference": ref_np, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if parity is not None:
            line["multi_gpu_parity"] = parity
        if scaling_diag is not None:
            line["scaling_breakdown"] = scaling_diag
        if one_gpu is not None:
            line["one_gpu_same_grid"] = one_gpu
        if extra:
            line["also"] = extra
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
