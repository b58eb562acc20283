#!/usr/bin/env python
"""Quick device-time check of the ops on one GPU (development aid; bench.py is the contract).

    python scripts/kbench.py [workload ...]      # default: bench_1M global_1deg
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from veros_b200 import isoneutral, synthetic  # noqa: E402
from veros_b200.state import IsoState  # noqa: E402


def timeit(fn, states, n=20):
    for w in range(3):
        fn(states[w % len(states)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n):
        fn(states[k % len(states)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name in sys.argv[1:] or ["bench_1M", "global_1deg", "global_4deg", "acc"]:
    st = synthetic.make_workload(name)
    cells = st["nx"] * st["ny"] * st["nz"]
    states = [IsoState.from_numpy(st, "cuda:0") for _ in range(2 if name != "bench_1M" else 3)]
    plans = {id(s): isoneutral.StepPlan(s) for s in states}
    t_step = timeit(lambda s: plans[id(s)](), states)
    if os.environ.get("VEROS_B200_MEGA_STATS"):
        plans[id(states[0])]()
        torch.cuda.synchronize()
        print(f"{name}: scheduling stats {isoneutral.step_stats(states[0])}")
    for pl in plans.values():
        pl.capture()
    t_graph = timeit(lambda s: plans[id(s)](), states)
    print(f"{name}: step via CUDA graph {t_graph:8.1f} us")
    t_pre = timeit(isoneutral.isoneutral_diffusion_pre, states)
    t_dT = timeit(lambda s: isoneutral.isoneutral_diffusion(s, s.variables.temp, True), states)
    print(f"{name}: step {t_step:8.1f} us  ({cells / t_step / 1e3:.2f} Gcell/s, "
          f"{cells * 276 / t_step / 1e3 / 6558.7 * 100:.1f}% of 276 B/cell roofline)   pre-op {t_pre:8.1f} us   diffusion-op {t_dT:8.1f} us")
    del states
    torch.cuda.empty_cache()
