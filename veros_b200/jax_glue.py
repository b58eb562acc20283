"""JAX / XLA side of the drop-in: registers the entry points of libveros_b200.so as XLA GPU custom-call
targets and rebinds Veros's isoneutral functions to them.

    import veros_b200.jax_glue as glue
    glue.install()          # after `import veros.core`, backend == "jax", device == "gpu"

NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no jax/jaxlib, so this
module is written against the reference's own use of the same JAX APIs (veros/core/special/tdma_.py:
PyCapsule "xla._CUSTOM_CALL_TARGET" :25-27 of tdma_cuda_.pyx, jax.ffi.register_ffi_target(...,
platform="CUDA", api_version=0) :34-37, jax._src.interpreters.mlir.custom_call :173-180, Primitive /
def_abstract_eval / register_lowering :184-193; reference pin jax==0.11.0).  The tests exercise the
same C symbols through ctypes instead (veros_b200/_lib.py); INTEGRATION.md shows the wiring.

Nothing here falls back to jnp: if the library or a GPU is missing, install() raises.
"""
import ctypes

from . import _lib

_CAPSULE_NAME = b"xla._CUSTOM_CALL_TARGET"
_registered = False

# XLA target name -> C symbol
TARGETS = {
    "veros_b200_solve_implicit_f64": "veros_b200_solve_implicit_f64",
    "veros_b200_iso_pre_f64": "veros_b200_iso_pre_f64",
    "veros_b200_iso_diffusion_f64": "veros_b200_iso_diffusion_f64",
    "veros_b200_iso_step_f64": "veros_b200_iso_step_f64",
    "veros_b200_vertmix_tempsalt_f64": "veros_b200_vertmix_tempsalt_f64",
    # neighbours of the path
    "veros_b200_implicit_vert_friction_f64": "veros_b200_implicit_vert_friction_f64",
    "veros_b200_advect_tempsalt_f64": "veros_b200_advect_tempsalt_f64",
    "veros_b200_set_eke_diffusivities_f64": "veros_b200_set_eke_diffusivities_f64",
    "veros_b200_iso_diag_streamfunction_f64": "veros_b200_iso_diag_streamfunction_f64",
    # the reference's own target names, served by the z-major compatible kernels (shim below)
    "tdma_cuda_double": "veros_b200_tdma_zmajor_f64",
    "tdma_cuda_float": "veros_b200_tdma_zmajor_f32",
}


def capsule(symbol):
    """PyCapsule around a C entry point, named as XLA expects (tdma_cuda_.pyx:23-27)."""
    fn = getattr(_lib.lib(), symbol)
    addr = ctypes.cast(fn, ctypes.c_void_p).value
    new = ctypes.pythonapi.PyCapsule_New
    new.restype = ctypes.py_object
    new.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
    return new(addr, _CAPSULE_NAME, None)


# ---- drop-in for veros.core.special.tdma_cuda_ (tdma_cuda_.pyx:18-30) ------------------------------
def build_tridiag_descriptor(num_systems, system_depth):
    return bytes(_lib.TridiagDescriptor(num_systems=int(num_systems), system_depth=int(system_depth)))


def gpu_custom_call_targets():
    return {name: capsule(TARGETS[name]) for name in ("tdma_cuda_double", "tdma_cuda_float")}


def install_tdma_shim():
    """Make the UNMODIFIED reference `veros.core.special.tdma_` use this library's TDMA: it imports
    `veros.core.special.tdma_cuda_` and reads `gpu_custom_call_targets` / `build_tridiag_descriptor`
    from it (tdma_.py:9-14,34-37,165).  Must run before `veros.core.special.tdma_` is imported."""
    import sys
    import types

    mod = types.ModuleType("veros.core.special.tdma_cuda_")
    mod.gpu_custom_call_targets = gpu_custom_call_targets()
    mod.build_tridiag_descriptor = build_tridiag_descriptor
    sys.modules["veros.core.special.tdma_cuda_"] = mod
    return mod


# ---- XLA registration -----------------------------------------------------------------------------
def register():
    global _registered
    if _registered:
        return
    import jax

    for name, symbol in TARGETS.items():
        if name.startswith("tdma_cuda_"):
            continue  # registered by the reference's tdma_.py through the shim
        jax.ffi.register_ffi_target(name, capsule(symbol), platform="CUDA", api_version=0)
    _registered = True


def _custom_call(name, operands, result_avals, descriptor, aliases):
    """Emit one custom call with native row-major layouts and in-place state (operand_output_aliases)."""
    import jaxlib.mlir.ir as ir
    from jax._src.interpreters.mlir import custom_call
    from jax.interpreters import mlir

    result_types = [ir.RankedTensorType.get(a.shape, mlir.dtype_to_ir_type(a.dtype)) for a in result_avals]
    return custom_call(
        name.encode(), operands=operands, result_types=result_types, backend_config=bytes(descriptor),
        operand_output_aliases=aliases,
    ).results


def _make_primitive(name, n_inout, inout_first, has_workspace, workspace_bytes, fresh_like=()):
    """Primitive whose first/last `n_inout` operands are returned updated (aliased), followed by fresh
    results shaped like the operands listed in `fresh_like`, followed by the scratch buffer."""
    import numpy as np
    from jax.core import ShapedArray
    from jax.extend.core import Primitive
    from jax.interpreters import mlir, xla

    prim = Primitive(name)
    prim.multiple_results = True

    def abstract_eval(*avals, descriptor):
        # the kernels are float64 / uint8 / int32 only: anything else would be read with the wrong element
        # size (the reference supports float32 runs, veros/runtime.py:93; this library does not)
        for a in avals:
            if np.dtype(a.dtype) not in (np.dtype(np.float64), np.dtype(np.uint8), np.dtype(np.int32)):
                raise TypeError(f"{name}: operand dtype {a.dtype} is not supported (float64 / uint8 / int32 only)")
        n = len(avals)
        idx = range(n_inout) if inout_first else range(n - n_inout, n)
        outs = [ShapedArray(avals[i].shape, avals[i].dtype) for i in idx]
        outs += [ShapedArray(avals[i].shape, avals[i].dtype) for i in fresh_like]
        if has_workspace:
            outs.append(ShapedArray((max(1, workspace_bytes(descriptor) // 8),), np.float64))
        return outs

    def lowering(ctx, *operands, descriptor):
        n = len(operands)
        idx = list(range(n_inout)) if inout_first else list(range(n - n_inout, n))
        aliases = {op: res for res, op in enumerate(idx)}
        return _custom_call(name, operands, ctx.avals_out, descriptor, aliases)

    prim.def_impl(lambda *a, **k: xla.apply_primitive(prim, *a, **k))
    prim.def_abstract_eval(abstract_eval)
    mlir.register_lowering(prim, lowering, platform="cuda")
    return prim


_prims = {}


def _primitives():
    if _prims:
        return _prims
    L = _lib.lib()

    def ws(fn):
        return lambda d: int(getattr(L, fn)(bytes(d), len(bytes(d))))

    _prims["pre"] = _make_primitive("veros_b200_iso_pre_f64", 7, False, True, ws("veros_b200_iso_pre_workspace_bytes"))
    _prims["diffusion"] = _make_primitive("veros_b200_iso_diffusion_f64", 3, True, True,
                                          ws("veros_b200_iso_diffusion_workspace_bytes"))
    _prims["step"] = _make_primitive("veros_b200_iso_step_f64", 12, True, True, ws("veros_b200_iso_step_workspace_bytes"))
    _prims["solve"] = _make_primitive("veros_b200_solve_implicit_f64", 0, True, False, None, fresh_like=(0,))
    # temp, salt updated in place; dtemp_vmix, dsalt_vmix fresh, shaped like kappaH (operand 3)
    _prims["vertmix"] = _make_primitive("veros_b200_vertmix_tempsalt_f64", 2, True, False, None, fresh_like=(3, 3))
    # u, v, du_mix, dv_mix, K_diss_v in place + scratch of 2 N M nz doubles
    _prims["friction"] = _make_primitive("veros_b200_implicit_vert_friction_f64", 5, True, True,
                                         lambda d: 16 * _lib.ColumnDescriptor.from_buffer_copy(bytes(d)).nx_tot *
                                         _lib.ColumnDescriptor.from_buffer_copy(bytes(d)).ny_tot *
                                         _lib.ColumnDescriptor.from_buffer_copy(bytes(d)).nz)
    _prims["advect"] = _make_primitive("veros_b200_advect_tempsalt_f64", 4, True, False, None)  # temp, salt, dtemp, dsalt
    _prims["streamfunction"] = _make_primitive("veros_b200_iso_diag_streamfunction_f64", 2, False, False, None)  # B1_gm, B2_gm
    return _prims


def _iso_descriptor(state, flags=0):
    st, vs = state.settings, state.variables
    N, M, nz = vs.K_iso.shape
    return _lib.IsoDescriptor(
        nx_tot=N, ny_tot=M, nz=nz, eq_of_state_type=st.eq_of_state_type,
        enable_conserve_energy=int(st.enable_conserve_energy), flags=flags, K_iso_steep=st.K_iso_steep,
        iso_slopec=st.iso_slopec, iso_dslope=st.iso_dslope, dt_tracer=st.dt_tracer, grav=st.grav, rho_0=st.rho_0)


_METRICS = ("dxt", "dxu", "dyt", "dyu", "cost", "cosu", "dzt", "dzw")
_PRE_OUT = ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33")


def _i32(x):
    import jax.numpy as jnp

    return jnp.asarray(x, dtype=jnp.int32).reshape((1,))


def _u8(x):
    import jax.numpy as jnp

    return x.astype(jnp.uint8)


def make_replacements():
    """Builds @veros_kernel / @veros_routine objects with the reference's signatures
    (isoneutral.py:18, diffusion.py:286-295, utilities.py:51-59)."""
    import jax.numpy as jnp
    from veros import KernelOutput, veros_kernel, veros_routine

    P = _primitives()

    @veros_kernel
    def isoneutral_diffusion_pre(state):
        vs = state.variables
        ops = [vs.temp, vs.salt, _i32(vs.tau), vs.K_iso, _u8(vs.maskT), _u8(vs.maskU), _u8(vs.maskV), _u8(vs.maskW)]
        ops += [getattr(vs, n) for n in _METRICS] + [vs.zt] + [getattr(vs, n) for n in _PRE_OUT]
        out = P["pre"].bind(*ops, descriptor=bytes(_iso_descriptor(state)))  # bytes: hashable jit-static param
        return KernelOutput(**dict(zip(_PRE_OUT, out[:7])))

    @veros_kernel(static_args=("istemp", "skew"))
    def _diffusion_kernel(state, tr, istemp, skew):
        vs, st = state.variables, state.settings
        energy = st.enable_conserve_energy
        dtr = vs.dtemp_iso if istemp else vs.dsalt_iso
        dummy = jnp.zeros((1,), dtype=tr.dtype)
        Pd = (vs.P_diss_skew if skew else vs.P_diss_iso) if energy else dummy
        X = (vs.int_drhodT if istemp else vs.int_drhodS) if energy else dummy
        K = vs.K_gm if skew else vs.K_iso
        ops = [tr, dtr, Pd, _i32(vs.tau), _i32(vs.taup1), K] + [getattr(vs, n) for n in _PRE_OUT]
        ops += [_u8(vs.maskT), _u8(vs.maskW), vs.kbot.astype(jnp.int32)] + [getattr(vs, n) for n in _METRICS] + [X]
        tr_new, dtr_new, P_new, _ = P["diffusion"].bind(
            *ops, descriptor=bytes(_iso_descriptor(state, _lib.FLAG_SKEW if skew else 0)))
        out = {("temp" if istemp else "salt"): tr_new, ("dtemp_iso" if istemp else "dsalt_iso"): dtr_new}
        if energy:
            out["P_diss_skew" if skew else "P_diss_iso"] = P_new
        return KernelOutput(**out)

    @veros_routine
    def isoneutral_diffusion(state, tr, istemp):
        state.variables.update(_diffusion_kernel(state, tr, istemp, False))

    @veros_routine
    def isoneutral_skew_diffusion(state, tr, istemp):
        state.variables.update(_diffusion_kernel(state, tr, istemp, True))

    @veros_kernel
    def solve_implicit(a, b, c, d, water_mask, edge_mask, b_edge=None, d_edge=None):
        if not a.shape == b.shape == c.shape == d.shape:
            raise ValueError("all inputs must have identical shape")
        if not a.dtype == b.dtype == c.dtype == d.dtype:
            raise ValueError("all inputs must have the same dtype")
        if a.dtype != jnp.float64:  # same error as veros_b200/utilities.py and tdma_.py:56-57
            raise TypeError(f"solve_tridiagonal only supports float64 arrays, got: {a.dtype}")
        nz = a.shape[-1]
        flags = (_lib.HAS_B_EDGE if b_edge is not None else 0) | (_lib.HAS_D_EDGE if d_edge is not None else 0)
        desc = _lib.SolveDescriptor(num_systems=a.size // nz, system_depth=nz, flags=flags, reserved=0)
        ops = [a, b, c, d, _u8(water_mask), _u8(edge_mask), b if b_edge is None else b_edge, d if d_edge is None else d_edge]
        return P["solve"].bind(*ops, descriptor=bytes(desc))[0]

    def solve_tridiagonal(a, b, c, d, water_mask, edge_mask):
        return solve_implicit(a, b, c, d, water_mask, edge_mask)

    @veros_kernel
    def vertmix_tempsalt(state):
        """veros/core/thermodynamics.py:248-300: one custom call, then the reference's own boundary treatment."""
        from veros.core import utilities
        from veros.core.operators import at, update

        vs, st = state.variables, state.settings
        N, M, nz = vs.kappaH.shape
        desc = _lib.VmixDescriptor(nx_tot=N, ny_tot=M, nz=nz, flags=0, dt_tracer=st.dt_tracer)
        ops = [vs.temp, vs.salt, _i32(vs.taup1), vs.kappaH, vs.forc_temp_surface, vs.forc_salt_surface,
               vs.kbot.astype(jnp.int32), vs.dzt, vs.dzw]
        temp, salt, dtemp_vmix, dsalt_vmix = P["vertmix"].bind(*ops, descriptor=bytes(desc))
        temp = update(temp, at[..., vs.taup1], utilities.enforce_boundaries(temp[..., vs.taup1], st.enable_cyclic_x))
        salt = update(salt, at[..., vs.taup1], utilities.enforce_boundaries(salt[..., vs.taup1], st.enable_cyclic_x))
        return KernelOutput(dtemp_vmix=dtemp_vmix, temp=temp, dsalt_vmix=dsalt_vmix, salt=salt)

    @veros_kernel
    def isoneutral_fused_step(state):
        """thermodynamics.py:430-432 as ONE custom call (veros_b200_iso_step_f64); returns all twelve arrays."""
        from .facade import STEP_OUTPUTS

        vs, st = state.variables, state.settings
        energy = st.enable_conserve_energy
        dummy = jnp.zeros((1,), dtype=vs.temp.dtype)
        inout = [vs.temp, vs.salt, vs.dtemp_iso, vs.dsalt_iso, vs.P_diss_iso if energy else dummy]
        inout += [getattr(vs, n) for n in _PRE_OUT]
        ops = inout + [_i32(vs.tau), _i32(vs.taup1), vs.K_iso, _u8(vs.maskT), _u8(vs.maskU), _u8(vs.maskV), _u8(vs.maskW),
                       vs.kbot.astype(jnp.int32)]
        ops += [getattr(vs, n) for n in _METRICS] + [vs.zt]
        ops += [vs.int_drhodT if energy else dummy, vs.int_drhodS if energy else dummy]
        out = P["step"].bind(*ops, descriptor=bytes(_iso_descriptor(state)))
        res = dict(zip(STEP_OUTPUTS, out[:12]))
        if not energy:
            res.pop("P_diss_iso")
        return KernelOutput(**res)

    @veros_kernel
    def implicit_vert_friction(state):
        """veros/core/friction.py:92-205 as one custom call (coefficient assembly fused into the column solves)."""
        vs, st = state.variables, state.settings
        N, M, nz = vs.kappaM.shape
        desc = _lib.ColumnDescriptor(nx_tot=N, ny_tot=M, nz=nz, flags=0, dt=st.dt_mom)
        ops = [vs.u, vs.v, vs.du_mix, vs.dv_mix, vs.K_diss_v, _i32(vs.tau), _i32(vs.taup1), vs.kappaM, _u8(vs.maskU),
               _u8(vs.maskV), vs.kbot.astype(jnp.int32), vs.dzt, vs.dzw, vs.dxt, vs.dxu, vs.area_v, vs.area_t]
        u, v, du_mix, dv_mix, K_diss_v, _ = P["friction"].bind(*ops, descriptor=bytes(desc))
        return KernelOutput(u=u, v=v, du_mix=du_mix, dv_mix=dv_mix, K_diss_v=K_diss_v)

    @veros_kernel
    def isoneutral_diag_streamfunction_kernel(state):
        """veros/core/isoneutral/isoneutral.py:232-258."""
        vs = state.variables
        N, M, nz = vs.K_gm.shape
        desc = _lib.ColumnDescriptor(nx_tot=N, ny_tot=M, nz=nz, flags=0, dt=0.0)
        B1, B2 = P["streamfunction"].bind(vs.K_gm, vs.Ai_ez, vs.Ai_nz, vs.B1_gm, vs.B2_gm, descriptor=bytes(desc))
        return KernelOutput(B1_gm=B1, B2_gm=B2)

    return dict(isoneutral_fused_step=isoneutral_fused_step, implicit_vert_friction=implicit_vert_friction,
                isoneutral_diag_streamfunction_kernel=isoneutral_diag_streamfunction_kernel,
                vertmix_tempsalt=vertmix_tempsalt, isoneutral_diffusion_pre=isoneutral_diffusion_pre, isoneutral_diffusion=isoneutral_diffusion,
                isoneutral_skew_diffusion=isoneutral_skew_diffusion, solve_implicit=solve_implicit,
                solve_tridiagonal=solve_tridiagonal)


def install(fused=True):
    """Idempotent; re-apply after anything reloads veros.core (test/pyom_consistency/conftest.py:21-32).

    The three functions of the call surface are rebound on `veros.core.isoneutral` (every caller sees them).  With
    `fused` (default) the model's own call site, veros/core/thermodynamics.py:430-432, additionally goes through
    `veros_b200.facade`: its `isoneutral` module global becomes a stand-in whose `isoneutral_diffusion_pre` is the ONE
    fused step op and whose `isoneutral_diffusion` has nothing left to do, so a model step runs the fused kernels
    (what bench.py times) instead of three separate custom calls."""
    from veros import runtime_settings as rs

    if rs.backend != "jax" or rs.device != "gpu":
        raise RuntimeError("veros_b200 has no CPU path: it needs backend='jax' and device='gpu'")
    if rs.float_type != "float64":
        raise RuntimeError(f"veros_b200 kernels are float64 only (runtime_settings.float_type = {rs.float_type!r})")
    _lib.lib()
    register()
    import veros.core.isoneutral as iso_pkg
    import veros.core.operators as operators
    import veros.core.utilities as utilities

    r = make_replacements()
    iso_pkg.isoneutral_diffusion_pre = r["isoneutral_diffusion_pre"]
    iso_pkg.isoneutral_diffusion = r["isoneutral_diffusion"]
    iso_pkg.isoneutral_skew_diffusion = r["isoneutral_skew_diffusion"]
    utilities.solve_implicit = r["solve_implicit"]
    utilities.solve_tridiagonal = r["solve_tridiagonal"]
    operators.solve_tridiagonal = r["solve_tridiagonal"]
    import veros.core.thermodynamics as thermodynamics

    thermodynamics.vertmix_tempsalt = r["vertmix_tempsalt"]  # looked up as a module global at :440
    import veros.core.friction as friction_mod
    import veros.core.isoneutral.isoneutral as iso_mod

    friction_mod.implicit_vert_friction = r["implicit_vert_friction"]  # called at friction.py:986-987
    iso_mod.isoneutral_diag_streamfunction_kernel = r["isoneutral_diag_streamfunction_kernel"]  # isoneutral.py:269
    from . import facade

    if fused:
        facade.install_facade(thermodynamics, iso_pkg, r["isoneutral_fused_step"])
    else:
        facade.uninstall_facade(thermodynamics, iso_pkg)
    return r
