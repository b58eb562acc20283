"""ctypes binding of libveros_b200.so (the C ABI declared in include/veros_b200.h).

There is no fallback: if the CUDA library has not been built, importing an op raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libveros_b200.so")

ABI_VERSION = 1

OPS = (
    "veros_b200_solve_implicit_f64",
    "veros_b200_tdma_zmajor_f64",
    "veros_b200_tdma_zmajor_f32",
    "veros_b200_iso_pre_f64",
    "veros_b200_iso_diffusion_f64",
    "veros_b200_iso_step_f64",
    "veros_b200_vertmix_tempsalt_f64",
    "veros_b200_implicit_vert_friction_f64",
    "veros_b200_iso_diag_streamfunction_f64",
    "veros_b200_set_eke_diffusivities_f64",
    "veros_b200_advect_tempsalt_f64",
)
HELPERS = (
    "veros_b200_iso_pre_workspace_bytes",
    "veros_b200_iso_diffusion_workspace_bytes",
    "veros_b200_iso_step_workspace_bytes",
    "veros_b200_iso_step_stats_offset",
    "veros_b200_last_error",
    "veros_b200_last_error_string",
    "veros_b200_clear_error",
    "veros_b200_abi_version",
    "veros_b200_descriptor_size",
    "veros_b200_launch_count",
    "veros_b200_profile_events",
    "veros_b200_halo_pack_unpack",
    "veros_b200_halo_put",
    "veros_b200_ipc_get_handle",
    "veros_b200_ipc_open_handle",
    "veros_b200_ipc_close",
)


class TridiagDescriptor(ctypes.Structure):
    """VerosB200TridiagDescriptor == the reference's TridiagDescriptor (cuda_tdma_kernels.h:5-8)."""

    _fields_ = [("num_systems", ctypes.c_int32), ("system_depth", ctypes.c_int32)]


class SolveDescriptor(ctypes.Structure):
    _fields_ = [("num_systems", ctypes.c_int32), ("system_depth", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class IsoDescriptor(ctypes.Structure):
    _fields_ = [
        ("nx_tot", ctypes.c_int32), ("ny_tot", ctypes.c_int32), ("nz", ctypes.c_int32),
        ("eq_of_state_type", ctypes.c_int32), ("enable_conserve_energy", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("K_iso_steep", ctypes.c_double), ("iso_slopec", ctypes.c_double), ("iso_dslope", ctypes.c_double),
        ("dt_tracer", ctypes.c_double), ("grav", ctypes.c_double), ("rho_0", ctypes.c_double),
    ]


class VmixDescriptor(ctypes.Structure):
    _fields_ = [("nx_tot", ctypes.c_int32), ("ny_tot", ctypes.c_int32), ("nz", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("dt_tracer", ctypes.c_double)]


class ColumnDescriptor(ctypes.Structure):
    _fields_ = [("nx_tot", ctypes.c_int32), ("ny_tot", ctypes.c_int32), ("nz", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("dt", ctypes.c_double)]


class EkeDescriptor(ctypes.Structure):
    _fields_ = [("nx_tot", ctypes.c_int32), ("ny_tot", ctypes.c_int32), ("nz", ctypes.c_int32),
                ("enable_eke", ctypes.c_int32), ("enable_eke_isopycnal_diffusion", ctypes.c_int32),
                ("flags", ctypes.c_int32)] + \
               [(k, ctypes.c_double) for k in ("pi", "eke_lmin", "eke_cross", "eke_crhin", "eke_k_max", "eke_c_k",
                                               "K_gm_0", "K_iso_0")]


class AdvectDescriptor(ctypes.Structure):
    _fields_ = [("nx_tot", ctypes.c_int32), ("ny_tot", ctypes.c_int32), ("nz", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("dt_tracer", ctypes.c_double), ("AB_eps", ctypes.c_double)]


ADVECT_SUPERBEE, ADVECT_NO_AB = 1, 2
HAS_B_EDGE, HAS_D_EDGE = 1, 2
FLAG_SKEW = 1
FLAG_NO_WEST_RING, FLAG_NO_EAST_RING = 2, 4
FLAG_PRE_SINGLE, FLAG_PRE_SPLIT = 8, 16
FLAG_STEP_FUSED = 32
FLAG_NO_MASK_SKIP = 64

_lib = None


def lib():
    """Load the library once.  Raises RuntimeError (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m veros_b200.build` "
            "(veros_b200 has no CPU or PyTorch fallback)")
    L = ctypes.CDLL(LIB_PATH)
    for name in OPS:
        fn = getattr(L, name)
        fn.restype = None
        fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_size_t]
    for name in ("veros_b200_iso_pre_workspace_bytes", "veros_b200_iso_diffusion_workspace_bytes",
                 "veros_b200_iso_step_workspace_bytes", "veros_b200_iso_step_stats_offset"):
        fn = getattr(L, name)
        fn.restype = ctypes.c_size_t
        fn.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    L.veros_b200_last_error.restype = ctypes.c_int
    L.veros_b200_last_error_string.restype = ctypes.c_char_p
    L.veros_b200_clear_error.restype = None
    L.veros_b200_abi_version.restype = ctypes.c_int
    L.veros_b200_descriptor_size.restype = ctypes.c_size_t
    L.veros_b200_descriptor_size.argtypes = [ctypes.c_int]
    L.veros_b200_launch_count.restype = ctypes.c_ulonglong
    L.veros_b200_halo_pack_unpack.restype = None
    L.veros_b200_halo_pack_unpack.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p]
    L.veros_b200_ipc_get_handle.restype = ctypes.c_int
    L.veros_b200_ipc_get_handle.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    L.veros_b200_ipc_open_handle.restype = ctypes.c_void_p
    L.veros_b200_ipc_open_handle.argtypes = [ctypes.c_int, ctypes.c_char_p]
    L.veros_b200_ipc_close.restype = None
    L.veros_b200_ipc_close.argtypes = [ctypes.c_void_p]
    L.veros_b200_halo_put.restype = None
    L.veros_b200_halo_put.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p)] + \
        [ctypes.c_int] * 7 + [ctypes.c_void_p] * 4
    L.veros_b200_profile_events.restype = None
    L.veros_b200_profile_events.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
    if L.veros_b200_abi_version() != ABI_VERSION:
        raise RuntimeError("libveros_b200.so ABI version mismatch; rebuild with `python -m veros_b200.build --force`")
    for which, cls in enumerate((TridiagDescriptor, SolveDescriptor, IsoDescriptor, VmixDescriptor, ColumnDescriptor,
                                 EkeDescriptor, AdvectDescriptor)):
        if L.veros_b200_descriptor_size(which) != ctypes.sizeof(cls):
            raise RuntimeError(f"descriptor layout mismatch for {cls.__name__}")
    _lib = L
    return L


def check_error(where=""):
    L = lib()
    code = L.veros_b200_last_error()
    if code:
        msg = L.veros_b200_last_error_string().decode()
        L.veros_b200_clear_error()
        raise RuntimeError(f"veros_b200 {where}: {msg}")


def call(symbol, buffers, descriptor, stream):
    """Invoke one custom-call symbol: `buffers` are raw device pointers (ints), operands then results."""
    L = lib()
    arr = (ctypes.c_void_p * len(buffers))(*buffers)
    opaque = bytes(descriptor)
    getattr(L, symbol)(ctypes.c_void_p(stream), arr, opaque, len(opaque))
    check_error(symbol)


def launch_count():
    return int(lib().veros_b200_launch_count())
