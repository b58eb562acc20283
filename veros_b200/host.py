"""Reference-facing entry point with HOST buffers: NumPy state in, NumPy state out.

This is the call a model running on host arrays makes per time step (the hot-path section of
veros/core/thermodynamics.py:425-432): the step's inputs are copied host -> device from pinned
staging buffers, the fused isoneutral step runs, and every array the step produces is copied back.
Static fields (masks, kbot, grid metrics) are uploaded once at construction.
"""
import numpy as np
import torch

from . import isoneutral
from .state import IsoState

STEP_INPUTS = ("temp", "salt", "K_iso", "dtemp_iso", "dsalt_iso", "P_diss_iso", "int_drhodT", "int_drhodS")
STEP_OUTPUTS = ("temp", "salt", "dtemp_iso", "dsalt_iso", "P_diss_iso",
                "Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33")


class HostStepper:
    def __init__(self, st, device="cuda"):
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

    def stage(self, host_arrays):
        """Copy the caller's NumPy arrays of this step into the pinned staging buffers."""
        for n, arr in host_arrays.items():
            self.pin_in[n].copy_(torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)))

    def step(self, synchronize=True):
        """H2D of the staged inputs, one fused isoneutral step, D2H of all outputs."""
        vs = self.state.variables
        for n in self.inputs:
            getattr(vs, n).copy_(self.pin_in[n], non_blocking=True)
        isoneutral.isoneutral_step(self.state)
        for n in self.outputs:
            self.pin_out[n].copy_(getattr(vs, n), non_blocking=True)
        if synchronize:
            torch.cuda.current_stream(self.state.device).synchronize()
        return {n: t.numpy() for n, t in self.pin_out.items()}
