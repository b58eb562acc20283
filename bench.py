#!/usr/bin/env python
"""Benchmark of the isoneutral step (pre + isoneutral_diffusion(temp) + (salt), column solves included).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]

One "step" = one pass of the hot path (veros/core/thermodynamics.py:430-432) over one synthetic
state; unit of work = one interior grid cell through one step ("cell-update").  Prints ONE JSON line
(rank 0).  N > 1 is launched by torch.distributed.run, one rank per GPU; every rank owns an x-slab
of the same size (weak scaling) and exchanges the 2-cell tracer halos with its ring neighbours over
NCCL after each step.  See DESIGN.md "Measurement" for how each field is obtained.
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
DEFAULT_WORKLOAD = "bench_1M"  # BASELINE.json configs[1]: isoneutral_benchmark.py, ~1M cells
L2_BYTES = 126e6


def algorithmic_bytes_per_cell(energy):
    """fused step 244 B/cell (energy off) / 276 B/cell (energy on, reference default)."""
    return 276 if energy else 244


PRE_BYTES_PER_CELL = 180  # isoneutral_diffusion_pre alone: 28 B read + 152 B written


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


def cpu_baseline(workload, seconds=12.0, threads=None):
    """The CPU oracle (oracle/iso_oracle.c, a port of the reference's NumPy path pinned to its golden
    vectors) timed on this host: repeated full steps of the same workload until `seconds` have passed."""
    from oracle import oracle
    from veros_b200 import synthetic

    if threads:
        oracle.set_num_threads(threads)
    cores = oracle.num_threads()
    name, kw = workload
    st = synthetic.make_workload(name, **kw)
    cells = st["nx"] * st["ny"] * st["nz"]
    oracle.isoneutral_step(st)  # warm-up (page faults, thread pool)
    n, t0 = 0, time.perf_counter()
    times = []
    while True:
        t1 = time.perf_counter()
        oracle.isoneutral_step(st)
        times.append(time.perf_counter() - t1)
        n += 1
        if time.perf_counter() - t0 > seconds or n >= 400:
            break
    best = min(times)
    return {
        "value": cells / best, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{n} full steps of {name} {st['nx']}x{st['ny']}x{st['nz']} ({cells} cells), best of {n}, "
                  f"C restatement of the reference NumPy path, OpenMP over x-planes",
        "mean_value": cells * n / sum(times),
    }, cells, min(times) * 1e3


def workload_kwargs(args, world):
    """Per-rank slab of the named workload (weak scaling: every rank gets the full named size)."""
    return args.workload, {}


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = workload_kwargs(args, 1)
    base, cells, ms = cpu_baseline(wl, seconds=max(5.0, min(60.0, 4.0 * args.steps)))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "note": "CPU oracle port of the reference NumPy path on host cores; "
                   "the Python reference itself cannot travel to the GPU box"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--replicas", type=int, default=0, help="state replicas rotated through (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the global_1deg side measurement")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (forced above 20 M cells per rank)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="N > 1: halo exchange by stores into the neighbours' memory over NVLink (one kernel per rank, CUDA "
                         "IPC mappings) or by pack + NCCL send/recv + unpack")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank gets a slab of the named size; strong: the named grid is split over the ranks")
    ap.add_argument("--profile", action="store_true",
                    help="only warm-up + timed steps + per-kernel pass (for ncu launch lists): no e2e, cpu or extra legs")
    ap.add_argument("--overlap", default="auto", choices=["auto", "on", "off"],
                    help="N > 1: hide the halo exchange behind interior compute (auto: slabs of >= 3 M cells; "
                         "below that the two extra boundary-strip passes cost more than the exchange)")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3)
    if args.profile:
        args.no_cpu = args.no_extra = True

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

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

    name, kw = workload_kwargs(args, world)
    if world > 1 and name != "bench_1M":
        nxl = synthetic.WORKLOADS[name]["nx"]
        if args.scaling == "strong":
            if nxl % world:
                raise SystemExit(f"--scaling strong: nx = {nxl} is not divisible by {world} ranks")
            nxl //= world
        kw = dict(nx=nxl, x_offset=rank * nxl, nx_global=nxl * world)
    elif world > 1:
        kw = dict(seed=17 + rank)
    st = synthetic.make_workload(name, **kw)
    nx, ny, nz = st["nx"], st["ny"], st["nz"]
    cells = nx * ny * nz
    energy = bool(st["enable_conserve_energy"])
    cyclic = bool(st.get("enable_cyclic_x", True))

    # state replicas: consecutive steps touch different memory, so nothing is served from L2
    probe = IsoState.from_numpy(st, dev)
    state_bytes = sum(t.numel() * t.element_size() for t in vars(probe.variables).values() if hasattr(t, 'numel'))
    replicas = args.replicas or max(2, min(8, int(3 * L2_BYTES / state_bytes) + 1))
    if state_bytes > 4e9 and not args.replicas:
        replicas = 1  # one pass over the state already streams several L2 sizes
    states = [probe] + [IsoState.from_numpy(st, dev) for _ in range(replicas - 1)]

    plans = [isoneutral.StepPlan(s) for s in states]  # argument marshalling done once, as under XLA
    overlap = world > 1 and (args.overlap == "on" or (args.overlap == "auto" and cells >= 3_000_000))
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

    def step(s):
        q = states.index(s)
        if world == 1:
            plans[q]()
        elif steppers is not None:
            steppers[q].step()  # boundary strips -> NCCL exchange || interior
        else:
            plans[q]()
            exchanges[q]()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for w in range(args.warmup):
        step(states[w % replicas])
    barrier()
    # ---- timed region: EXACTLY K steps -------------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    barrier()
    e0.record()
    for k in range(args.steps):
        step(states[k % replicas])
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * cells * args.steps / (ms_total * 1e-3)

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
    kern_ms = {
        "setup (+eos5)": sum(r[0].elapsed_time(r[1]) for r in evs) / ksteps,
        "iso_pre_kernel (slopes + tensor + fluxes)": sum(r[1].elapsed_time(r[2]) for r in evs) / ksteps,
        "update_kernel (divergence + column solve + tendencies + dissipation)": sum(r[2].elapsed_time(r[3]) for r in evs) / ksteps,
    }
    t_k1 = kern_ms["iso_pre_kernel (slopes + tensor + fluxes)"]

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
        s_.settings.enable_cyclic_x = False  # kernel only: the exchange is timed with the step above
    vev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(ksteps)]
    thermodynamics.vertmix_tempsalt(states[0])  # the reference-facing call once (allocates the tendencies)
    for s_ in states[1:]:
        s_.variables.dtemp_vmix, s_.variables.dsalt_vmix = states[0].variables.dtemp_vmix, states[0].variables.dsalt_vmix
    vplans = [thermodynamics.VertmixPlan(s_) for s_ in states]
    for k in range(ksteps):
        vev[k][0].record()
        vplans[k % replicas]()
        vev[k][1].record()
    torch.cuda.synchronize()
    vmix_ms = sum(e[0].elapsed_time(e[1]) for e in vev) / ksteps
    VMIX_BYTES = 56  # R temp,salt@taup1 16 + kappaH 8; W temp,salt 16 + dtemp_vmix,dsalt_vmix 16

    peak, peak_src = load_peaks()
    step_bytes = algorithmic_bytes_per_cell(energy)
    step_gbs = cells * step_bytes / (ms_step * 1e-3) / 1e9
    # dominant kernel of the step; algorithmic bytes: isoneutral_diffusion_pre 180 B/cell, the rest of the fused
    # step (276 or 244 B/cell) belongs to the update kernel
    k_pre = "iso_pre_kernel (slopes + tensor + fluxes)"
    k_upd = "update_kernel (divergence + column solve + tendencies + dissipation)"
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
    if args.profile:
        if rank == 0:
            sampler.stop()
            emit({"profile_only": True, "ms_per_step": ms_step, "kernels_ms": kern_ms})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end: host buffers in, host buffers out ----------------------------------------------
    del states[1:]
    torch.cuda.empty_cache()
    hs = None
    if args.no_e2e or cells > 20_000_000:
        clocks = sampler.stop() if rank == 0 else None
        e2e = None  # the pinned staging buffers of this leg would be tens of GB
    else:
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
        e2e = {"value": world * cells * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": hs.h2d_bytes,
               "d2h_bytes_per_step": hs.d2h_bytes, "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps}

    extra = None
    if rank == 0 and world == 1 and not args.no_extra and name != "global_1deg":
        # the north-star grid (1 degree, 6.6 M cells, EOS 5) measured the same way, for context
        del hs
        torch.cuda.empty_cache()
        st1 = synthetic.make_workload("global_1deg")
        c1 = st1["nx"] * st1["ny"] * st1["nz"]
        ss = [IsoState.from_numpy(st1, dev) for _ in range(2)]
        for w in range(3):
            isoneutral.isoneutral_step(ss[w % 2])
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n1 = 20
        a0.record()
        for k in range(n1):
            isoneutral.isoneutral_step(ss[k % 2])
        a1.record()
        torch.cuda.synchronize()
        ms1 = a0.elapsed_time(a1) / n1
        gbs1 = c1 * step_bytes / (ms1 * 1e-3) / 1e9
        extra = {"workload": "global_1deg 360x160x115 analytic, EOS 5", "value": c1 / (ms1 * 1e-3), "ms_per_step": ms1,
                 "step_roofline_frac": gbs1 / peak, "achieved_gbs": gbs1}
        del ss

    base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        base, _, _ = cpu_baseline((name, kw))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": name, "nx": nx, "ny": ny, "nz": nz, "cells_per_gpu": cells,
                "eq_of_state_type": int(st["eq_of_state_type"]), "enable_conserve_energy": energy,
                "parallelism": f"x-slabs x{world}" + (((" + ring halo exchange of temp/salt[taup1] by peer-memory stores over NVLink"
                                                        if args.halo == "peer" else " + NCCL ring halo exchange of temp/salt[taup1]") +
                                (", boundary strips first, exchange overlapped with interior" if overlap else ", exchange after the step") + halo_note)
                                if world > 1 else ""),
                "l2": (f"inputs larger than L2: {replicas} state replicas of {state_bytes / 1e6:.0f} MB rotated, "
                       f"no replica is touched twice in a row") if replicas > 1 else
                      f"inputs larger than L2: one state of {state_bytes / 1e6:.0f} MB streamed per step",
            },
            "roofline": roofline, "step_roofline": step_roofline, "next_ops": next_ops, "cpu_baseline": base, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if extra:
            line["also"] = extra
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
