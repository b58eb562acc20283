"""Implicit vertical friction: same name, argument and return convention as the reference kernel
``veros.core.friction.implicit_vert_friction`` (veros/core/friction.py:92-205)

    vs.update(friction.implicit_vert_friction(state))

-- the first other caller of ``solve_implicit`` with its coefficient assembly fused
into the column solve (csrc/next_ops.cu).  Bit-identical to the reference's NumPy backend.
"""
import torch

from . import _lib
from .state import KernelOutput

NEEDS = ("u", "v", "du_mix", "dv_mix", "K_diss_v", "tau", "taup1", "kappaM", "maskU", "maskV", "kbot", "dzt", "dzw",
         "dxt", "dxu", "area_v", "area_t")


def implicit_vert_friction(state):
    vs, settings = state.variables, state.settings
    for name in NEEDS:
        if getattr(vs, name, None) is None:
            raise ValueError(f"implicit_vert_friction needs variable {name}")
    if not vs.u.is_cuda:
        raise RuntimeError("veros_b200 has no CPU path: the state must live on a CUDA device")
    N, M, nz = settings.nx + 4, settings.ny + 4, settings.nz
    if tuple(vs.kappaM.shape) != (N, M, nz) or tuple(vs.u.shape) != (N, M, nz, 3) or tuple(vs.area_v.shape) != (N, M):
        raise ValueError("u / kappaM / area_v do not match the grid")
    desc = _lib.ColumnDescriptor(nx_tot=N, ny_tot=M, nz=nz, flags=0, dt=float(settings.dt_mom))
    ws = state.workspace(2 * N * M * nz * 8)
    inout = [vs.u, vs.v, vs.du_mix, vs.dv_mix, vs.K_diss_v]
    operands = inout + [vs.tau, vs.taup1, vs.kappaM, vs.maskU, vs.maskV, vs.kbot, vs.dzt, vs.dzw, vs.dxt, vs.dxu,
                        vs.area_v, vs.area_t]
    _lib.call("veros_b200_implicit_vert_friction_f64", [int(t.data_ptr()) for t in operands + inout + [ws]], desc,
              torch.cuda.current_stream(state.device).cuda_stream)
    return KernelOutput(u=vs.u, v=vs.v, du_mix=vs.du_mix, dv_mix=vs.dv_mix, K_diss_v=vs.K_diss_v)
