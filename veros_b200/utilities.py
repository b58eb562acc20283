"""Column-solve op: same signatures as ``veros.core.utilities.solve_implicit`` (utilities.py:51-59)
and ``veros.core.operators.solve_tridiagonal`` (operators.py:60-77, 103-156), on CUDA tensors in the
model's native (X, Y, nz) layout.  Error behaviour follows veros/core/special/tdma_.py:53-57.
"""
import torch

from . import _lib


def _check(a, b, c, d, water_mask, edge_mask):
    if not a.shape == b.shape == c.shape == d.shape:
        raise ValueError("all inputs must have identical shape")
    if not a.dtype == b.dtype == c.dtype == d.dtype:
        raise ValueError("all inputs must have the same dtype")
    if a.dtype != torch.float64:
        raise TypeError(f"solve_tridiagonal only supports float64 arrays, got: {a.dtype}")
    if water_mask.shape != a.shape or edge_mask.shape != a.shape:
        raise ValueError("masks must have the shape of the diagonals")
    if not a.is_cuda:
        raise RuntimeError("veros_b200 has no CPU path: inputs must be CUDA tensors")


def _as_u8(m):
    if m.dtype == torch.bool:
        return m.contiguous().view(torch.uint8)
    if m.dtype != torch.uint8:
        raise TypeError("masks must be bool or uint8")
    return m.contiguous()


def solve_implicit(a, b, c, d, water_mask, edge_mask, b_edge=None, d_edge=None):
    _check(a, b, c, d, water_mask, edge_mask)
    a, b, c, d = (t.contiguous() for t in (a, b, c, d))
    water_mask, edge_mask = _as_u8(water_mask), _as_u8(edge_mask)
    out = torch.empty_like(a)
    nz = a.shape[-1] if a.dim() else 0
    ncol = a.numel() // nz if nz else 0
    if ncol == 0:
        return out
    flags = 0
    if b_edge is not None:
        b_edge = b_edge.contiguous()
        flags |= _lib.HAS_B_EDGE
    if d_edge is not None:
        d_edge = d_edge.contiguous()
        flags |= _lib.HAS_D_EDGE
    desc = _lib.SolveDescriptor(num_systems=ncol, system_depth=nz, flags=flags, reserved=0)
    bufs = [a, b, c, d, water_mask, edge_mask, b_edge if b_edge is not None else b,
            d_edge if d_edge is not None else d, out]
    _lib.call("veros_b200_solve_implicit_f64", [int(t.data_ptr()) for t in bufs], desc,
              torch.cuda.current_stream(a.device).cuda_stream)
    return out


def solve_tridiagonal(a, b, c, d, water_mask, edge_mask):
    return solve_implicit(a, b, c, d, water_mask, edge_mask)


def tdma_zmajor(a, b, c, d):
    """The reference's own custom-call contract (tdma_.py:129-181): operands already masked
    (tdma_.py:63-66) and in the z-major layout XLA is asked for; `a..d` are (X, Y, nz) tensors whose
    *memory* is laid out [z][x][y] (i.e. ``t.permute(2, 0, 1).contiguous().permute(1, 2, 0)``)."""
    if not a.shape == b.shape == c.shape == d.shape:
        raise ValueError("all inputs must have identical shape")
    if not a.dtype == b.dtype == c.dtype == d.dtype:
        raise ValueError("all inputs must have the same dtype")
    if a.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"TDMA only supports float32/float64 arrays, got: {a.dtype}")
    nz = a.shape[-1]
    nsys = a.numel() // nz
    zm = [t.permute(2, 0, 1) for t in (a, b, c, d)]
    if not all(t.is_contiguous() for t in zm):
        raise ValueError("operands must be z-major in memory")
    out = torch.empty_like(zm[0])
    work = torch.empty_like(zm[0])
    desc = _lib.TridiagDescriptor(num_systems=nsys, system_depth=nz)
    sym = "veros_b200_tdma_zmajor_f64" if a.dtype == torch.float64 else "veros_b200_tdma_zmajor_f32"
    _lib.call(sym, [int(t.data_ptr()) for t in zm + [out, work]], desc,
              torch.cuda.current_stream(a.device).cuda_stream)
    return out.permute(1, 2, 0)
