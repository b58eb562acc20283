#!/usr/bin/env python
"""Experiment (round 2): does running the step x-block by x-block keep the inter-kernel scratch (fluxes, TEOS-10
derivatives, staged int_drhod*) in the 126 MB L2?  Uses the existing bit-exact sub-slab mechanism (IsoState.subslab):
every block recomputes 4 halo planes, so the block version does MORE work; a lower time means the L2 effect wins.

    python scripts/exp_l2block.py [workload]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from veros_b200 import isoneutral, synthetic  # noqa: E402
from veros_b200.state import IsoState  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "global_1deg"
st = synthetic.make_workload(name)
cells = st["nx"] * st["ny"] * st["nz"]
states = [IsoState.from_numpy(st, "cuda:0") for _ in range(2)]
N = st["nx"] + 4
for width in (0, 120, 60, 30, 20, 12):
    plans = []
    for s in states:
        s.workspace(isoneutral.step_workspace_bytes(s))
        if width == 0:
            plans.append([isoneutral.StepPlan(s)])
            continue
        cuts = list(range(2, N - 2, width)) + [N - 2]
        row = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            row.append(isoneutral.StepPlan(s.subslab(0 if a == 2 else a - 2, N if b == N - 2 else b + 2)))
        plans.append(row)

    def step(k):
        for p in plans[k % 2]:
            p()

    for w in range(3):
        step(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for k in range(n):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name}: x-blocks of {width or N - 4:4d} planes ({len(plans[0]):3d} blocks): {ms:7.3f} ms/step  "
          f"{cells * 276 / ms / 1e6 / 6558.7 * 100:5.1f} % of the 276 B/cell roofline", flush=True)
