"""Host side of the isoneutral mixing step: same names, arguments and return conventions as
``veros.core.isoneutral`` (veros/core/isoneutral/__init__.py), executing hand-written sm_100a
kernels through the XLA-custom-call C ABI of libveros_b200.so.

    isoneutral_diffusion_pre(state) -> KernelOutput(Ai_ez, Ai_nz, Ai_bx, Ai_by, K_11, K_22, K_33)
        veros/core/isoneutral/isoneutral.py:18-229
    isoneutral_diffusion(state, tr, istemp) -> None (updates temp|salt, dtemp_iso|dsalt_iso, P_diss_iso)
        veros/core/isoneutral/diffusion.py:286-295
    isoneutral_skew_diffusion(state, tr, istemp) -> None
        veros/core/isoneutral/diffusion.py:298-307
    isoneutral_step(state) -> None: the three calls of veros/core/thermodynamics.py:430-432 as one op

All work is enqueued on the current CUDA stream of ``state.device``; nothing synchronises.
"""
import torch

from . import _lib
from .state import KernelOutput

_PRE_OUT = ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33")
_METRICS = ("dxt", "dxu", "dyt", "dyu", "cost", "cosu", "dzt", "dzw")


def _descriptor(state, flags=0):
    st = state.settings
    return _lib.IsoDescriptor(
        nx_tot=st.nx + 4, ny_tot=st.ny + 4, nz=st.nz, eq_of_state_type=st.eq_of_state_type,
        enable_conserve_energy=int(st.enable_conserve_energy), flags=flags | getattr(state, "ring_flags", 0) | getattr(state, "tuning_flags", 0),
        K_iso_steep=st.K_iso_steep, iso_slopec=st.iso_slopec, iso_dslope=st.iso_dslope,
        dt_tracer=st.dt_tracer, grav=st.grav, rho_0=st.rho_0)


def _stream(state):
    return torch.cuda.current_stream(state.device).cuda_stream


def _ptrs(tensors):
    return [int(t.data_ptr()) for t in tensors]


def isoneutral_diffusion_pre(state):
    """Isopycnal slopes and mixing tensor; returns the seven updated arrays (updated in place)."""
    vs = state.variables
    desc = _descriptor(state)
    opaque = bytes(desc)
    ws = state.workspace(_lib.lib().veros_b200_iso_pre_workspace_bytes(opaque, len(opaque)))
    inout = [getattr(vs, n) for n in _PRE_OUT]
    operands = [vs.temp, vs.salt, vs.tau, vs.K_iso, vs.maskT, vs.maskU, vs.maskV, vs.maskW]
    operands += [getattr(vs, n) for n in _METRICS] + [vs.zt] + inout
    _lib.call("veros_b200_iso_pre_f64", _ptrs(operands + inout + [ws]), desc, _stream(state))
    return KernelOutput(**dict(zip(_PRE_OUT, inout)))


def _diffusion(state, tr, istemp, skew):
    vs, st = state.variables, state.settings
    if tr is not vs.temp and tr is not vs.salt:
        if tr.data_ptr() not in (vs.temp.data_ptr(), vs.salt.data_ptr()):
            raise ValueError("tr must be state.variables.temp or state.variables.salt")
    energy = st.enable_conserve_energy
    dtracer = vs.dtemp_iso if istemp else vs.dsalt_iso
    dummy = state.dummy()
    if energy:
        P = vs.P_diss_skew if skew else vs.P_diss_iso
        X = vs.int_drhodT if istemp else vs.int_drhodS
    else:
        P = X = dummy
    K = vs.K_gm if skew else vs.K_iso
    desc = _descriptor(state, _lib.FLAG_SKEW if skew else 0)
    opaque = bytes(desc)
    ws = state.workspace(_lib.lib().veros_b200_iso_diffusion_workspace_bytes(opaque, len(opaque)))
    operands = [tr, dtracer, P, vs.tau, vs.taup1, K, vs.Ai_ez, vs.Ai_nz, vs.Ai_bx, vs.Ai_by, vs.K_11, vs.K_22,
                vs.K_33, vs.maskT, vs.maskW, vs.kbot] + [getattr(vs, n) for n in _METRICS] + [X]
    results = [tr, dtracer, P, ws]
    _lib.call("veros_b200_iso_diffusion_f64", _ptrs(operands + results), desc, _stream(state))


def isoneutral_diffusion(state, tr, istemp):
    """Isopycnal diffusion of one tracer incl. the implicit K_33 column solve (in place on state)."""
    _diffusion(state, tr, istemp, skew=False)


def isoneutral_skew_diffusion(state, tr, istemp):
    """GM skew diffusion of one tracer (in place on state)."""
    _diffusion(state, tr, istemp, skew=True)


def isoneutral_diag_streamfunction(state):
    """Horizontal components of the eddy-driven streamfunction from K_gm and the Ai_ez / Ai_nz the path produced
    (veros/core/isoneutral/isoneutral.py:232-269); returns KernelOutput(B1_gm, B2_gm) like the reference kernel."""
    vs, st = state.variables, state.settings
    for name in ("K_gm", "Ai_ez", "Ai_nz", "B1_gm", "B2_gm"):
        if getattr(vs, name, None) is None:
            raise ValueError(f"isoneutral_diag_streamfunction needs variable {name}")
    desc = _lib.ColumnDescriptor(nx_tot=st.nx + 4, ny_tot=st.ny + 4, nz=st.nz, flags=0, dt=0.0)
    operands = [vs.K_gm, vs.Ai_ez, vs.Ai_nz, vs.B1_gm, vs.B2_gm]
    _lib.call("veros_b200_iso_diag_streamfunction_f64", _ptrs(operands + [vs.B1_gm, vs.B2_gm]), desc, _stream(state))
    return KernelOutput(B1_gm=vs.B1_gm, B2_gm=vs.B2_gm)


def step_workspace_bytes(state):
    """Scratch the fused step needs for `state` (veros_b200_iso_step_workspace_bytes)."""
    opaque = bytes(_descriptor(state))
    return int(_lib.lib().veros_b200_iso_step_workspace_bytes(opaque, len(opaque)))


def step_stats(state):
    """Scheduling statistics of the last fused step on `state` (development aid): per item type the time its CTAs
    spent waiting for dependencies / working (summed over CTAs, microseconds) and the number of items."""
    opaque = bytes(_descriptor(state))
    off = int(_lib.lib().veros_b200_iso_step_stats_offset(opaque, len(opaque)))
    ws = state.workspace(_lib.lib().veros_b200_iso_step_workspace_bytes(opaque, len(opaque)))
    raw = ws.view(torch.int64)[off // 8: off // 8 + 16].cpu().tolist()
    names = {1: "eos", 2: "copy", 3: "pre", 4: "update"}
    return {n: dict(wait_us=raw[t] / 1e3, work_us=raw[5 + t] / 1e3, items=raw[10 + t]) for t, n in names.items()}


class StepPlan:
    """`isoneutral_step(state)` with the argument marshalling done once.

    Under JAX the custom call is invoked by XLA with no Python in between; this is the equivalent for
    the ctypes harness: buffer list, descriptor and workspace are built at construction (the state's
    tensors must not be reallocated afterwards), a call is one foreign-function call on the current
    stream.  Matters for small grids, where marshalling ~45 pointers costs more than the kernels."""

    def __init__(self, state):
        import ctypes

        self.state = state
        ptrs, desc, keep = _step_arguments(state)
        self._keep = keep  # keeps the scratch tensors alive
        self._arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        self._opaque = bytes(desc)
        self._fn = _lib.lib().veros_b200_iso_step_f64
        self._err = _lib.lib().veros_b200_last_error
        self._void_p = ctypes.c_void_p

        self._graph = None

    def __call__(self):
        if self._graph is not None:
            self._graph.replay()
            return
        self._fn(self._void_p(torch.cuda.current_stream(self.state.device).cuda_stream), self._arr, self._opaque,
                 len(self._opaque))
        if self._err():
            _lib.check_error("veros_b200_iso_step_f64")

    def capture(self):
        """Record the step's kernel sequence in a CUDA graph (one launch per step afterwards): removes
        the launch gaps that dominate small grids (4 degree: 54 k cells).  The ops only enqueue work,
        so they are capture-safe once each kernel has run once (first launches set function attributes)."""
        self()  # warm-up outside capture
        torch.cuda.current_stream(self.state.device).synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._fn(self._void_p(torch.cuda.current_stream(self.state.device).cuda_stream), self._arr, self._opaque,
                     len(self._opaque))
        if self._err():
            _lib.check_error("veros_b200_iso_step_f64 (graph capture)")
        self._graph = g
        return self


def isoneutral_step(state):
    """pre + isoneutral_diffusion(temp) + isoneutral_diffusion(salt) (thermodynamics.py:430-432)."""
    ptrs, desc, _ = _step_arguments(state)
    _lib.call("veros_b200_iso_step_f64", ptrs, desc, _stream(state))


def _step_arguments(state):
    vs, st = state.variables, state.settings
    energy = st.enable_conserve_energy
    dummy = state.dummy()
    desc = _descriptor(state)
    opaque = bytes(desc)
    ws = state.workspace(_lib.lib().veros_b200_iso_step_workspace_bytes(opaque, len(opaque)))
    inout = [vs.temp, vs.salt, vs.dtemp_iso, vs.dsalt_iso, vs.P_diss_iso if energy else dummy]
    inout += [getattr(vs, n) for n in _PRE_OUT]
    operands = inout + [vs.tau, vs.taup1, vs.K_iso, vs.maskT, vs.maskU, vs.maskV, vs.maskW, vs.kbot]
    operands += [getattr(vs, n) for n in _METRICS] + [vs.zt]
    operands += [vs.int_drhodT if energy else dummy, vs.int_drhodS if energy else dummy]
    return _ptrs(operands + inout + [ws]), desc, (ws, dummy)
