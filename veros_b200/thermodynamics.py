"""The step right after the isoneutral path: ``vertmix_tempsalt`` (veros/core/thermodynamics.py:248-300).

Same name, argument and return convention as the reference kernel:

    vs.update(thermodynamics.vertmix_tempsalt(state))

One launch of ``veros_b200_vertmix_tempsalt_f64`` (csrc/vertmix.cu: coefficient assembly from kappaH,
both column solves on one dgtsv factorisation, the dtemp_vmix / dsalt_vmix tendencies), followed by
the boundary treatment of :290-297 -- ``enforce_boundaries`` is the x-halo exchange of
``veros_b200.decomp`` (cyclic wrap on one process, NCCL ring between slabs).
"""
import torch

from . import _lib, decomp
from .state import KernelOutput

NEEDS = ("temp", "salt", "taup1", "kappaH", "forc_temp_surface", "forc_salt_surface", "kbot", "dzt", "dzw")


def _arguments(state):
    vs, settings = state.variables, state.settings
    for name in NEEDS:
        if getattr(vs, name, None) is None:
            raise ValueError(f"vertmix_tempsalt needs variable {name}")
    N, M, nz = settings.nx + 4, settings.ny + 4, settings.nz
    if tuple(vs.kappaH.shape) != (N, M, nz) or tuple(vs.forc_temp_surface.shape) != (N, M):
        raise ValueError("kappaH / forc_temp_surface do not match the grid")
    if not vs.temp.is_cuda:
        raise RuntimeError("veros_b200 has no CPU path: the state must live on a CUDA device")
    for name in ("dtemp_vmix", "dsalt_vmix"):
        if getattr(vs, name, None) is None:
            setattr(vs, name, torch.empty((N, M, nz), dtype=torch.float64, device=state.device))
    desc = _lib.VmixDescriptor(nx_tot=N, ny_tot=M, nz=nz, flags=0, dt_tracer=float(settings.dt_tracer))
    operands = [vs.temp, vs.salt, vs.taup1, vs.kappaH, vs.forc_temp_surface, vs.forc_salt_surface, vs.kbot,
                vs.dzt, vs.dzw]
    results = [vs.temp, vs.salt, vs.dtemp_vmix, vs.dsalt_vmix]  # operand_output_aliases {0: 0, 1: 1}
    return [int(t.data_ptr()) for t in operands + results], desc


def vertmix_tempsalt(state, group=None):
    vs, settings = state.variables, state.settings
    ptrs, desc = _arguments(state)
    _lib.call("veros_b200_vertmix_tempsalt_f64", ptrs, desc, torch.cuda.current_stream(state.device).cuda_stream)
    decomp.exchange_halos_x([vs.temp, vs.salt], cyclic=bool(getattr(settings, "enable_cyclic_x", False)),
                            group=group, level=vs.taup1_host)
    return KernelOutput(dtemp_vmix=vs.dtemp_vmix, temp=vs.temp, dsalt_vmix=vs.dsalt_vmix, salt=vs.salt)


ADVECT_NEEDS = ("temp", "salt", "dtemp", "dsalt", "tau", "taup1", "u", "v", "w", "maskT", "maskU", "maskV", "maskW", "dxt", "dyt",
                "dzt", "cost", "cosu")


def advect_tempsalt(state, adams_bashforth=True):
    """advect_temperature + advect_salinity (veros/core/thermodynamics.py:43-62, i.e. d{temp,salt}[..., tau] =
    advect_tracer(...), :10-40) and, unless `adams_bashforth` is False, the Adams-Bashforth step of :223-245 that
    produces temp/salt[..., taup1] -- one kernel launch for both tracers (csrc/advect.cu), bit-identical to the
    reference's NumPy backend.  Returns KernelOutput(temp, salt, dtemp, dsalt) like the reference's kernels."""
    vs, settings = state.variables, state.settings
    for name in ADVECT_NEEDS:
        if getattr(vs, name, None) is None:
            raise ValueError(f"advect_tempsalt needs variable {name}")
    if not vs.temp.is_cuda:
        raise RuntimeError("veros_b200 has no CPU path: the state must live on a CUDA device")
    N, M, nz = settings.nx + 4, settings.ny + 4, settings.nz
    if tuple(vs.dtemp.shape) != (N, M, nz, 3) or tuple(vs.w.shape) != (N, M, nz, 3):
        raise ValueError("dtemp / w do not match the grid")
    if getattr(vs, "taum1", None) is None:  # taum1 = the third time level
        vs.taum1 = (3 - vs.tau - vs.taup1).to(torch.int32)
    flags = (_lib.ADVECT_SUPERBEE if getattr(settings, "enable_superbee_advection", False) else 0) | \
            (0 if adams_bashforth else _lib.ADVECT_NO_AB)
    desc = _lib.AdvectDescriptor(nx_tot=N, ny_tot=M, nz=nz, flags=flags, dt_tracer=float(settings.dt_tracer),
                                 AB_eps=float(getattr(settings, "AB_eps", 0.1)))
    inout = [vs.temp, vs.salt, vs.dtemp, vs.dsalt]
    operands = inout + [vs.tau, vs.taup1, vs.taum1, vs.u, vs.v, vs.w, vs.maskT, vs.maskU, vs.maskV, vs.maskW, vs.dxt, vs.dyt,
                        vs.dzt, vs.cost, vs.cosu]
    _lib.call("veros_b200_advect_tempsalt_f64", [int(t.data_ptr()) for t in operands + inout], desc,
              torch.cuda.current_stream(state.device).cuda_stream)
    return KernelOutput(temp=vs.temp, salt=vs.salt, dtemp=vs.dtemp, dsalt=vs.dsalt)


class VertmixPlan:
    """The kernel of `vertmix_tempsalt(state)` with the argument marshalling done once (cf.
    isoneutral.StepPlan): one foreign-function call per invocation, no boundary exchange."""

    def __init__(self, state):
        import ctypes

        self.state = state
        ptrs, desc = _arguments(state)
        self._arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        self._opaque = bytes(desc)
        self._fn = _lib.lib().veros_b200_vertmix_tempsalt_f64
        self._err = _lib.lib().veros_b200_last_error
        self._void_p = ctypes.c_void_p

    def __call__(self):
        self._fn(self._void_p(torch.cuda.current_stream(self.state.device).cuda_stream), self._arr, self._opaque,
                 len(self._opaque))
        if self._err():
            _lib.check_error("veros_b200_vertmix_tempsalt_f64")
