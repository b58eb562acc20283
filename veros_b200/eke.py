"""Producer of the diffusivities the isoneutral path reads: same name, argument and return convention as the
reference kernel ``veros.core.eke.set_eke_diffusivities_kernel`` (veros/core/eke.py:34-85)

    vs.update(eke.set_eke_diffusivities_kernel(state))

(SURVEY.md section 8f, rank 4).  One kernel (csrc/next_ops.cu); bit-identical to the reference's NumPy backend,
NumPy's pairwise order of the Rossby-radius column sum included.
"""
import torch

from . import _lib
from .state import KernelOutput


def set_eke_diffusivities_kernel(state):
    vs, st = state.variables, state.settings
    N, M, nz = st.nx + 4, st.ny + 4, st.nz
    on = bool(st.enable_eke)
    if getattr(vs, "K_gm", None) is None or getattr(vs, "K_iso", None) is None:
        raise ValueError("set_eke_diffusivities_kernel needs variables K_gm and K_iso")
    if not vs.K_gm.is_cuda:
        raise RuntimeError("veros_b200 has no CPU path: the state must live on a CUDA device")
    dev = state.device
    if on:
        for name in ("Nsqr", "eke", "tau", "maskW", "dzw", "coriolis_t", "beta"):
            if getattr(vs, name, None) is None:
                raise ValueError(f"set_eke_diffusivities_kernel needs variable {name}")
        for name, shape in (("L_rossby", (N, M)), ("L_rhines", (N, M, nz)), ("eke_len", (N, M, nz)), ("sqrteke", (N, M, nz))):
            if getattr(vs, name, None) is None:
                setattr(vs, name, torch.zeros(shape, dtype=torch.float64, device=dev))
        operands = [vs.Nsqr, vs.eke, vs.tau, vs.maskW, vs.dzw, vs.coriolis_t, vs.beta]
        results = [vs.L_rossby, vs.L_rhines, vs.eke_len, vs.sqrteke, vs.K_gm, vs.K_iso]
    else:
        d = state.dummy()
        operands = [d] * 7
        results = [d, d, d, d, vs.K_gm, vs.K_iso]
    desc = _lib.EkeDescriptor(
        nx_tot=N, ny_tot=M, nz=nz, enable_eke=int(on),
        enable_eke_isopycnal_diffusion=int(bool(getattr(st, "enable_eke_isopycnal_diffusion", False))), flags=0,
        pi=float(getattr(st, "pi", 3.14159265358979323846264338327950588)), eke_lmin=float(getattr(st, "eke_lmin", 0.0)),
        eke_cross=float(getattr(st, "eke_cross", 0.0)), eke_crhin=float(getattr(st, "eke_crhin", 0.0)),
        eke_k_max=float(getattr(st, "eke_k_max", 0.0)), eke_c_k=float(getattr(st, "eke_c_k", 0.0)),
        K_gm_0=float(st.K_gm_0), K_iso_0=float(st.K_iso_0))
    _lib.call("veros_b200_set_eke_diffusivities_f64", [int(t.data_ptr()) for t in operands + results], desc,
              torch.cuda.current_stream(dev).cuda_stream)
    if not on:
        return KernelOutput(K_gm=vs.K_gm, K_iso=vs.K_iso)
    return KernelOutput(L_rossby=vs.L_rossby, L_rhines=vs.L_rhines, eke_len=vs.eke_len, sqrteke=vs.sqrteke,
                        K_gm=vs.K_gm, K_iso=vs.K_iso)
