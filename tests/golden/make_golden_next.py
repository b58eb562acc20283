#!/usr/bin/env python
"""Golden vectors for the SURVEY.md 8(f) rank 3 / rank 4 neighbours of the hot path, from the REFERENCE's NumPy
implementation (runs only in the build container: imports veros from /root/reference; the ``fric_*.npz`` /
``sf_*.npz`` / ``eke_*.npz`` files it writes are committed).

    python tests/golden/make_golden_next.py

* ``fric_*``  implicit_vert_friction (veros/core/friction.py:92-205): inputs u, v, kappaM, maskU, maskV, kbot,
  K_diss_v, du_mix, dv_mix, dzt, dzw, dxt, dxu, area_v, area_t, tau, taup1, dt_mom; outputs u, v, du_mix, dv_mix,
  K_diss_v.  Random kappaM makes dgtsv interchange rows in many columns.
* ``sf_*``    isoneutral_diag_streamfunction_kernel (veros/core/isoneutral/isoneutral.py:232-258): inputs K_gm,
  Ai_ez, Ai_nz, B1_gm, B2_gm; outputs B1_gm, B2_gm.
* ``adv_*``   advect_tracer (veros/core/thermodynamics.py:10-40, superbee and 2nd-order fluxes of advection.py),
  advect_temperature / advect_salinity and the Adams-Bashforth step (:223-245).
* ``eke_*``   set_eke_diffusivities_kernel (veros/core/eke.py:34-85), both branches (enable_eke on / off).
"""
import os
import sys

os.environ.setdefault("VEROS_BACKEND", "numpy")
os.environ.setdefault("VEROS_LOGLEVEL", "error")
os.environ.setdefault("VEROS_DISKLESS_MODE", "1")
sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("VEROS_REFERENCE", "/root/reference"))

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, out):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e3:.0f} kB")


def random_state(seed, **extra):
    from veros.pyom_compat import get_random_state

    settings = dict(dt_tracer=3600, dt_mom=3600, enable_streamfunction=False, enable_conserve_energy=True,
                    enable_neutral_diffusion=True, enable_skew_diffusion=True)
    settings.update(extra)
    np.random.seed(seed)
    return get_random_state(extra_settings=settings)


def friction_case(name, seed, **extra):
    from veros.core import friction

    state = random_state(seed, **extra)
    vs, st = state.variables, state.settings
    out = {}
    for k in ("u", "v", "kappaM", "maskU", "maskV", "kbot", "K_diss_v", "du_mix", "dv_mix", "dzt", "dzw", "dxt", "dxu",
              "area_v", "area_t"):
        out["in__" + k] = np.array(getattr(vs, k))
    out["in__tau"], out["in__taup1"] = np.int32(vs.tau), np.int32(vs.taup1)
    for k in ("nx", "ny", "nz", "dt_mom"):
        out["set__" + k] = np.asarray(getattr(st, k))
    vs.update(friction.implicit_vert_friction(state))
    for k in ("u", "v", "du_mix", "dv_mix", "K_diss_v"):
        out["out__" + k] = np.array(getattr(vs, k))
    save(name, out)


def streamfunction_case(name, seed, **extra):
    from veros.core.isoneutral import isoneutral

    state = random_state(seed, **extra)
    vs, st = state.variables, state.settings
    out = {}
    for k in ("K_gm", "Ai_ez", "Ai_nz", "B1_gm", "B2_gm"):
        out["in__" + k] = np.array(getattr(vs, k))
    for k in ("nx", "ny", "nz"):
        out["set__" + k] = np.asarray(getattr(st, k))
    vs.update(isoneutral.isoneutral_diag_streamfunction_kernel(state))
    for k in ("B1_gm", "B2_gm"):
        out["out__" + k] = np.array(getattr(vs, k))
    save(name, out)


def eke_case(name, seed, **extra):
    from veros.core import eke

    state = random_state(seed, **extra)
    vs, st = state.variables, state.settings
    out = {}
    for k in ("Nsqr", "eke", "maskW", "dzw", "coriolis_t", "beta", "K_gm", "K_iso", "L_rossby", "L_rhines", "eke_len", "sqrteke"):
        try:
            out["in__" + k] = np.array(getattr(vs, k))
        except RuntimeError:  # EKE variables are inactive when enable_eke is off
            pass
    out["in__tau"] = np.int32(vs.tau)
    for k in ("nx", "ny", "nz", "enable_eke", "enable_eke_isopycnal_diffusion", "pi", "eke_lmin", "eke_cross", "eke_crhin",
              "eke_k_max", "eke_c_k", "K_gm_0", "K_iso_0"):
        out["set__" + k] = np.asarray(getattr(st, k))
    res = eke.set_eke_diffusivities_kernel(state)
    for k, v in res._asdict().items():
        out["out__" + k] = np.array(v)
    save(name, out)


def advection_case(name, seed, **extra):
    """advect_temperature + advect_salinity (thermodynamics.py:43-62 -> advect_tracer :10-40) and the Adams-Bashforth
    step of :223-245 (re-stated here from the reference's own expressions, which sit at the end of
    advect_temp_salt_enthalpy behind the energy diagnostics)."""
    from veros.core import thermodynamics

    state = random_state(seed, **extra)
    vs, st = state.variables, state.settings
    out = {}
    for k in ("temp", "salt", "dtemp", "dsalt", "u", "v", "w", "maskT", "maskU", "maskV", "maskW", "dxt", "dyt", "dzt",
              "cost", "cosu"):
        out["in__" + k] = np.array(getattr(vs, k))
    for k in ("tau", "taup1", "taum1"):
        out["in__" + k] = np.int32(getattr(vs, k))
    for k in ("nx", "ny", "nz", "dt_tracer", "AB_eps", "enable_superbee_advection"):
        out["set__" + k] = np.asarray(getattr(st, k))
    out["out__dtr_temp"] = np.array(thermodynamics.advect_tracer(state, vs.temp[..., vs.tau]))
    vs.update(thermodynamics.advect_temperature(state))
    vs.update(thermodynamics.advect_salinity(state))
    for tr, d in (("temp", "dtemp"), ("salt", "dsalt")):
        x, dx = getattr(vs, tr), getattr(vs, d)
        x[:, :, :, vs.taup1] = (x[:, :, :, vs.tau] + st.dt_tracer * ((1.5 + st.AB_eps) * dx[:, :, :, vs.tau]
                                                                    - (0.5 + st.AB_eps) * dx[:, :, :, vs.taum1]) * vs.maskT)
        out["out__" + tr], out["out__" + d] = np.array(x), np.array(dx)
    save(name, out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["fric", "sf", "eke", "adv"]
    if "adv" in which:
        advection_case("adv_superbee_10x8x7", 71, nx=10, ny=8, nz=7, enable_superbee_advection=True)
        advection_case("adv_superbee_6x5x2_cyclic", 72, nx=6, ny=5, nz=2, enable_superbee_advection=True, enable_cyclic_x=True)
        advection_case("adv_2nd_9x7x5", 73, nx=9, ny=7, nz=5, enable_superbee_advection=False)
    if "fric" in which:
        friction_case("fric_rand_10x8x7", 41, nx=10, ny=8, nz=7)
        friction_case("fric_rand_6x5x12_cyclic", 42, nx=6, ny=5, nz=12, enable_cyclic_x=True)
        friction_case("fric_rand_9x7x2", 43, nx=9, ny=7, nz=2)
    if "sf" in which:
        streamfunction_case("sf_rand_10x8x7", 51, nx=10, ny=8, nz=7)
        streamfunction_case("sf_rand_7x9x2", 52, nx=7, ny=9, nz=2)
    if "eke" in which:
        eke_case("eke_rand_10x8x9", 61, nx=10, ny=8, nz=9, enable_eke=True, enable_eke_isopycnal_diffusion=True)
        eke_case("eke_rand_6x7x140", 62, nx=6, ny=7, nz=140, enable_eke=True, enable_eke_isopycnal_diffusion=False)
        eke_case("eke_off_5x6x4", 63, nx=5, ny=6, nz=4, enable_eke=False)
