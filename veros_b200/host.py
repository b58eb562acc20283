"""Reference-facing entry point with HOST buffers: NumPy state in, NumPy state out.

This is the call a model running on host arrays makes per time step (the hot-path section of
veros/core/thermodynamics.py:425-432): the step's inputs are copied host -> device from pinned
staging buffers, the fused isoneutral step runs, and every array the step produces is copied back.
Static fields (masks, kbot, grid metrics) are uploaded once at construction.

The step is PCIe bound (375 MB per step for 1 M cells against 0.26 ms of kernels), so it is
pipelined over x-sub-slabs: while slab s computes, slab s+1 is on its way in and slab s-1 on its way
out, each on its own stream (the two copy directions use separate DMA engines).  Sub-slabs reproduce
the whole-slab result bit for bit (include/veros_b200.h, VEROS_B200_FLAG_NO_*_RING;
tests/test_gpu_parity.py::test_subslab_composition_is_bitexact).
"""
import numpy as np
import torch

from . import isoneutral
from .state import IsoState

STEP_INPUTS = ("temp", "salt", "K_iso", "dtemp_iso", "dsalt_iso", "P_diss_iso", "int_drhodT", "int_drhodS")
STEP_OUTPUTS = ("temp", "salt", "dtemp_iso", "dsalt_iso", "P_diss_iso",
                "Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33")


class HostStepper:
    def __init__(self, st, device="cuda", slabs=8):
        self.state = IsoState.from_numpy(st, device)
        vs = self.state.variables
        energy = self.state.settings.enable_conserve_energy
        skip = () if energy else ("P_diss_iso", "int_drhodT", "int_drhodS")
        self.inputs = [n for n in STEP_INPUTS if n not in skip]
        self.outputs = [n for n in STEP_OUTPUTS if n not in skip]
        self.pin_in = {n: torch.empty(getattr(vs, n).shape, dtype=torch.float64).pin_memory() for n in self.inputs}
        self.pin_out = {n: torch.empty(getattr(vs, n).shape, dtype=torch.float64).pin_memory() for n in self.outputs}
        for n in self.inputs:
            self.pin_in[n].copy_(torch.from_numpy(np.ascontiguousarray(st[n], dtype=np.float64)))
        self.h2d_bytes = sum(t.numel() * 8 for t in self.pin_in.values())
        self.d2h_bytes = sum(t.numel() * 8 for t in self.pin_out.values())

        # interior planes [2, N-2) split into `slabs` owned ranges [a, b); each computes on the view
        # [a-2, b+2) (clamped to the slab), is fed by the planes up to b+2 and returns its owned planes
        N = self.state.settings.nx + 4
        # sub-slabs share the parent's scratch: size it once for the largest user so that the pointers
        # captured by the plans stay valid
        self.state.workspace(isoneutral.step_workspace_bytes(self.state))
        slabs = max(1, min(int(slabs), N - 4))
        cuts = [2 + (N - 4) * s // slabs for s in range(slabs + 1)]
        self.parts = []
        fed = 0
        for s in range(slabs):
            a, b = cuts[s], cuts[s + 1]
            lo, hi = (0 if s == 0 else a - 2), (N if s == slabs - 1 else b + 2)
            own_lo, own_hi = (0 if s == 0 else a), (N if s == slabs - 1 else b)
            sub = self.state if slabs == 1 else self.state.subslab(lo, hi)
            self.parts.append(dict(sub=sub, plan=isoneutral.StepPlan(sub), feed=(fed, hi), own=(own_lo, own_hi),
                                   ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event()))
            fed = hi
        dev = self.state.device
        self.s_in = torch.cuda.Stream(dev)
        self.s_out = torch.cuda.Stream(dev)

    def stage(self, host_arrays):
        """Copy the caller's NumPy arrays of this step into the pinned staging buffers."""
        for n, arr in host_arrays.items():
            self.pin_in[n].copy_(torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)))

    def step(self, synchronize=True):
        """H2D of the staged inputs, one fused isoneutral step, D2H of all outputs -- slab-pipelined."""
        vs = self.state.variables
        cur = torch.cuda.current_stream(self.state.device)
        self.s_in.wait_stream(cur)
        self.s_out.wait_stream(cur)
        with torch.cuda.stream(self.s_in):
            for part in self.parts:
                lo, hi = part["feed"]
                for n in self.inputs:
                    getattr(vs, n)[lo:hi].copy_(self.pin_in[n][lo:hi], non_blocking=True)
                part["ev_in"].record(self.s_in)
        for part in self.parts:
            cur.wait_event(part["ev_in"])
            part["plan"]()
            part["ev_done"].record(cur)
            lo, hi = part["own"]
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(part["ev_done"])
                for n in self.outputs:
                    self.pin_out[n][lo:hi].copy_(getattr(vs, n)[lo:hi], non_blocking=True)
        cur.wait_stream(self.s_out)
        if synchronize:
            cur.synchronize()
        return {n: t.numpy() for n, t in self.pin_out.items()}
