"""ctypes front-end of oracle/iso_oracle.c (TEST INFRASTRUCTURE ONLY, see that file's header).

The functions mirror the reference's call surface on plain NumPy dictionaries ("state dicts" keyed by
the reference's variable names, veros/variables.py) so parity tests read like the reference's:

    oracle.isoneutral_diffusion_pre(st)          # veros/core/isoneutral/isoneutral.py:18
    oracle.isoneutral_diffusion(st, "temp")      # veros/core/isoneutral/diffusion.py:286
    oracle.isoneutral_skew_diffusion(st, "salt") # veros/core/isoneutral/diffusion.py:298
    oracle.solve_implicit(a, b, c, d, water, edge, b_edge=None, d_edge=None)  # veros/core/utilities.py:51
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleParams(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int32), ("M", ctypes.c_int32), ("nz", ctypes.c_int32),
        ("eos_type", ctypes.c_int32), ("enable_conserve_energy", ctypes.c_int32),
        ("tau", ctypes.c_int32), ("taup1", ctypes.c_int32), ("pad_", ctypes.c_int32),
        ("K_iso_steep", ctypes.c_double), ("iso_slopec", ctypes.c_double), ("iso_dslope", ctypes.c_double),
        ("dt_tracer", ctypes.c_double), ("grav", ctypes.c_double), ("rho_0", ctypes.c_double),
    ]


def build(force=False):
    """Compile liboracle.so next to the source (gcc, a second or two)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "iso_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def _p(arr):
    return None if arr is None else arr.ctypes.data_as(ctypes.c_void_p)


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _u8(x):
    return np.ascontiguousarray(np.asarray(x).astype(np.uint8))


def params_from_state(st):
    N, M, nz = st["K_iso"].shape
    return OracleParams(
        N=N, M=M, nz=nz, eos_type=int(st["eq_of_state_type"]),
        enable_conserve_energy=int(bool(st["enable_conserve_energy"])),
        tau=int(st["tau"]), taup1=int(st["taup1"]), pad_=0,
        K_iso_steep=float(st["K_iso_steep"]), iso_slopec=float(st["iso_slopec"]),
        iso_dslope=float(st["iso_dslope"]), dt_tracer=float(st["dt_tracer"]),
        grav=float(st["grav"]), rho_0=float(st["rho_0"]),
    )


_METRICS = ("dxt", "dxu", "dyt", "dyu", "cost", "cosu", "dzt", "dzw")


def isoneutral_diffusion_pre(st):
    """In place on st["Ai_*"], st["K_11|K_22|K_33"] (only the reference's write regions change)."""
    P = params_from_state(st)
    for k in ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33"):
        st[k] = _f64(st[k]).copy()
    args = [_f64(st["temp"]), _f64(st["salt"]), _f64(st["K_iso"])]
    args += [_u8(st[m]) for m in ("maskT", "maskU", "maskV", "maskW")]
    args += [_f64(st[m]) for m in _METRICS] + [_f64(st["zt"])]
    outs = [st[k] for k in ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33")]
    lib().oracle_iso_pre(ctypes.byref(P), *[_p(a) for a in args + outs])
    return st


def _diffusion(st, tracer, iso, tdma_mode=0):
    P = params_from_state(st)
    istemp = tracer == "temp"
    dname = "dtemp_iso" if istemp else "dsalt_iso"
    pname = "P_diss_iso" if iso else "P_diss_skew"
    xname = "int_drhodT" if istemp else "int_drhodS"
    st[tracer] = _f64(st[tracer]).copy()
    st[dname] = _f64(st[dname]).copy()
    energy = bool(st["enable_conserve_energy"])
    if energy:
        st[pname] = _f64(st[pname]).copy()
    shape = st["K_iso"].shape
    fe, fn, ft = (np.zeros(shape) for _ in range(3))
    kfield = _f64(st["K_iso"] if iso else st["K_gm"])
    args = [st[tracer], st[dname], kfield] + [_f64(st[k]) for k in ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33")]
    args += [_u8(st["maskT"]), _u8(st["maskW"]), np.ascontiguousarray(st["kbot"], dtype=np.int32)]
    args += [_f64(st[m]) for m in _METRICS]
    args += [_f64(st[xname]) if energy else None, st[pname] if energy else None, fe, fn, ft]
    lib().oracle_iso_diffusion(ctypes.byref(P), ctypes.c_int32(int(iso)), *[_p(a) for a in args],
                               ctypes.c_int32(tdma_mode))
    st["flux_east"], st["flux_north"], st["flux_top"] = fe, fn, ft
    return st


def isoneutral_diffusion(st, tracer, tdma_mode=0):
    return _diffusion(st, tracer, True, tdma_mode)


def isoneutral_skew_diffusion(st, tracer):
    return _diffusion(st, tracer, False)


def isoneutral_step(st, tdma_mode=0):
    """thermodynamics.py:430-432: pre, then T, then S."""
    isoneutral_diffusion_pre(st)
    isoneutral_diffusion(st, "temp", tdma_mode)
    isoneutral_diffusion(st, "salt", tdma_mode)
    return st


def vertmix_tempsalt(st, tdma_mode=0):
    """veros/core/thermodynamics.py:248-300 in place on st["temp"|"salt"] (time level taup1); creates
    st["dtemp_vmix"|"dsalt_vmix"].  Includes the single-process enforce_boundaries of :290-297."""
    N, M, nz = st["kappaH"].shape
    st["temp"] = _f64(st["temp"]).copy()
    st["salt"] = _f64(st["salt"]).copy()
    st["dtemp_vmix"] = np.zeros((N, M, nz))
    st["dsalt_vmix"] = np.zeros((N, M, nz))
    args = [st["temp"], st["salt"], _f64(st["kappaH"]), _f64(st["forc_temp_surface"]), _f64(st["forc_salt_surface"]),
            np.ascontiguousarray(st["kbot"], dtype=np.int32), _f64(st["dzt"]), _f64(st["dzw"]),
            st["dtemp_vmix"], st["dsalt_vmix"]]
    lib().oracle_vertmix_tempsalt(
        ctypes.c_int32(N), ctypes.c_int32(M), ctypes.c_int32(nz), ctypes.c_int32(int(st["taup1"])),
        ctypes.c_int32(int(bool(st.get("enable_cyclic_x", False)))), ctypes.c_double(float(st["dt_tracer"])),
        *[_p(a) for a in args], ctypes.c_int32(tdma_mode))
    return st


def solve_implicit(a, b, c, d, water_mask, edge_mask, b_edge=None, d_edge=None, mode=0):
    """mode 0: dgtsv (no-pivot) operation order; mode 1: Thomas cp/dp recurrence."""
    a, b, c, d = map(_f64, (a, b, c, d))
    nz = a.shape[-1]
    ncol = a.size // nz if nz else 0
    out = np.zeros(a.shape)
    if a.size == 0:
        return out
    be = None if b_edge is None else _f64(b_edge)
    de = None if d_edge is None else _f64(d_edge)
    lib().oracle_solve_implicit(ctypes.c_int64(ncol), ctypes.c_int32(nz), _p(a), _p(b), _p(c), _p(d),
                                _p(_u8(water_mask)), _p(_u8(edge_mask)), _p(be), _p(de), _p(out),
                                ctypes.c_int32(mode))
    return out


def solve_tridiagonal(a, b, c, d, water_mask, edge_mask, mode=0):
    return solve_implicit(a, b, c, d, water_mask, edge_mask, mode=mode)


def implicit_vert_friction(st, tdma_mode=0):
    """veros/core/friction.py:92-205 in place on st["u"|"v"] (time level taup1), st["du_mix"|"dv_mix"|"K_diss_v"]."""
    N, M, nz = st["kappaM"].shape
    for k in ("u", "v", "du_mix", "dv_mix", "K_diss_v"):
        st[k] = _f64(st[k]).copy()
    args = [st["u"], st["v"], _f64(st["kappaM"]), _u8(st["maskU"]), _u8(st["maskV"]),
            np.ascontiguousarray(st["kbot"], dtype=np.int32)]
    args += [_f64(st[k]) for k in ("dzt", "dzw", "dxt", "dxu", "area_v", "area_t")]
    args += [st["du_mix"], st["dv_mix"], st["K_diss_v"]]
    lib().oracle_implicit_vert_friction(
        ctypes.c_int32(N), ctypes.c_int32(M), ctypes.c_int32(nz), ctypes.c_int32(int(st["tau"])),
        ctypes.c_int32(int(st["taup1"])), ctypes.c_double(float(st["dt_mom"])), *[_p(a) for a in args],
        ctypes.c_int32(tdma_mode))
    return st


def isoneutral_diag_streamfunction(st):
    """veros/core/isoneutral/isoneutral.py:232-258 in place on st["B1_gm"|"B2_gm"]."""
    N, M, nz = st["K_gm"].shape
    st["B1_gm"], st["B2_gm"] = _f64(st["B1_gm"]).copy(), _f64(st["B2_gm"]).copy()
    lib().oracle_diag_streamfunction(ctypes.c_int32(N), ctypes.c_int32(M), ctypes.c_int32(nz), _p(_f64(st["K_gm"])),
                                     _p(_f64(st["Ai_ez"])), _p(_f64(st["Ai_nz"])), _p(st["B1_gm"]), _p(st["B2_gm"]))
    return st


def set_eke_diffusivities(st, sum_variant=1):
    """veros/core/eke.py:34-85; returns the dict of arrays the reference's KernelOutput carries."""
    N, M, nz = st["K_gm"].shape
    on = bool(st["enable_eke"])
    out = {k: np.zeros((N, M, nz)) for k in ("L_rhines", "eke_len", "sqrteke", "K_gm", "K_iso")}
    out["L_rossby"] = np.zeros((N, M))
    dummy3, dummy2 = np.zeros((N, M, nz, 3)), np.zeros((N, M))
    g = lambda k, d: _f64(st[k]) if (on and k in st) else d
    lib().oracle_set_eke_diffusivities(
        ctypes.c_int32(N), ctypes.c_int32(M), ctypes.c_int32(nz), ctypes.c_int32(int(st.get("tau", 0))),
        ctypes.c_int32(int(on)), ctypes.c_int32(int(bool(st["enable_eke_isopycnal_diffusion"]))),
        *[ctypes.c_double(float(st[k])) for k in ("pi", "eke_lmin", "eke_cross", "eke_crhin", "eke_k_max", "eke_c_k", "K_gm_0", "K_iso_0")],
        _p(g("Nsqr", dummy3)), _p(g("eke", dummy3)), _p(_u8(st["maskW"]) if "maskW" in st else np.zeros((N, M, nz), np.uint8)),
        _p(g("dzw", np.zeros(nz))), _p(g("coriolis_t", dummy2)), _p(g("beta", dummy2)),
        _p(out["L_rossby"]), _p(out["L_rhines"]), _p(out["eke_len"]), _p(out["sqrteke"]), _p(out["K_gm"]), _p(out["K_iso"]),
        ctypes.c_int32(sum_variant))
    if not on:
        return {"K_gm": out["K_gm"], "K_iso": out["K_iso"]}
    return out


def advect_tracer(st, tracer, lev=None):
    """thermodynamics.advect_tracer(state, st[tracer][..., lev]) (veros/core/thermodynamics.py:10-40); returns dtr."""
    N, M, nz = st["maskT"].shape
    lev = int(st["tau"]) if lev is None else int(lev)
    dtr = np.zeros((N, M, nz))
    args = [_f64(st[tracer]), _f64(st["u"]), _f64(st["v"]), _f64(st["w"])] + [_u8(st[m]) for m in ("maskT", "maskU", "maskV", "maskW")]
    args += [_f64(st[m]) for m in ("dxt", "dyt", "dzt", "cost", "cosu")] + [dtr]
    lib().oracle_advect_tracer(ctypes.c_int32(N), ctypes.c_int32(M), ctypes.c_int32(nz),
                               ctypes.c_int32(int(bool(st["enable_superbee_advection"]))), ctypes.c_int32(int(st["tau"])),
                               ctypes.c_int32(lev), ctypes.c_double(float(st["dt_tracer"])), *[_p(a) for a in args])
    return dtr


def advect_tempsalt(st):
    """advect_temperature, advect_salinity (thermodynamics.py:43-62) and the Adams-Bashforth step (:223-245), in place
    on st["dtemp"|"dsalt"][..., tau] and st["temp"|"salt"][..., taup1]."""
    N, M, nz = st["maskT"].shape
    tau, taup1 = int(st["tau"]), int(st["taup1"])
    taum1 = int(st.get("taum1", 3 - tau - taup1))
    for tr, d in (("temp", "dtemp"), ("salt", "dsalt")):
        st[tr], st[d] = _f64(st[tr]).copy(), _f64(st[d]).copy()
        st[d][..., tau] = advect_tracer(st, tr)
        lib().oracle_adams_bashforth(ctypes.c_int32(N), ctypes.c_int32(M), ctypes.c_int32(nz), ctypes.c_int32(tau),
                                     ctypes.c_int32(taup1), ctypes.c_int32(taum1), ctypes.c_double(float(st["dt_tracer"])),
                                     ctypes.c_double(float(st["AB_eps"])), _p(st[tr]), _p(st[d]), _p(_u8(st["maskT"])))
    return st
