#!/usr/bin/env python
"""Generate golden input/output vectors from the REFERENCE's own NumPy implementation.

Runs only in the build container (it imports veros from /root/reference, which does not exist on
the GPU box); the resulting ``tests/golden/*.npz`` files are committed and are what the tests read.

    python tests/golden/make_golden.py            # (re)writes every fixture

Each fixture holds, for one state:
  * every input of the hot path before the step,
  * the outputs of ``isoneutral_diffusion_pre`` (``pre__*``),
  * the outputs of ``isoneutral_diffusion(temp)``, then ``(salt)`` (``dT__*``, ``dS__*``), chained
    exactly as veros/core/thermodynamics.py:425-432 does,
  * the outputs of ``isoneutral_skew_diffusion`` for T and S (``kT__*``, ``kS__*``; :434-437),
  * the a,b,c,d / masks / result of the ``solve_tridiagonal`` call made inside the T solve
    (``tdma__*``; veros/core/isoneutral/diffusion.py:165 -> utilities.py:51-59 -> operators.py:60-77),
  * the fluxes returned by ``isoneutral_diffusion_tracer`` for T (``fluxT__*``).
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

INPUT_VARS = [
    "temp", "salt", "K_iso", "K_gm", "maskT", "maskU", "maskV", "maskW", "kbot",
    "dxt", "dxu", "dyt", "dyu", "cost", "cosu", "dzt", "dzw", "zt",
    "Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33",
    "dtemp_iso", "dsalt_iso", "P_diss_iso", "P_diss_skew", "int_drhodT", "int_drhodS",
]
SETTINGS = [
    "nx", "ny", "nz", "eq_of_state_type", "enable_conserve_energy", "enable_cyclic_x",
    "K_iso_steep", "iso_slopec", "iso_dslope", "dt_tracer", "grav", "rho_0",
]


def snapshot_inputs(state, out):
    vs, st = state.variables, state.settings
    for name in INPUT_VARS:
        if hasattr(vs, name):
            out["in__" + name] = np.array(getattr(vs, name))
    for name in ("tau", "taup1", "taum1"):
        out["in__" + name] = np.int32(getattr(vs, name))
    for name in SETTINGS:
        out["set__" + name] = np.asarray(getattr(st, name))


def run_hot_path(state, out):
    """thermodynamics.py:425-437 on `state`, recording every stage."""
    from veros.core import isoneutral, utilities
    from veros.core.isoneutral import diffusion as isodiff

    vs = state.variables

    pre = isoneutral.isoneutral_diffusion_pre(state)
    vs.update(pre)
    for name in ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33"):
        out["pre__" + name] = np.array(getattr(vs, name))

    # fluxes of the T solve (pure function of the state after `pre`)
    _, _, fe, fn, ft = isodiff.isoneutral_diffusion_tracer(state, vs.temp, vs.dtemp_iso, iso=True, skew=False)
    out["fluxT__flux_east"], out["fluxT__flux_north"], out["fluxT__flux_top"] = map(np.array, (fe, fn, ft))

    # capture the column solve inside the T step
    captured = {}
    orig = utilities.solve_tridiagonal

    def spy(a, b, c, d, water_mask, edge_mask):
        res = orig(a, b, c, d, water_mask, edge_mask)
        if not captured:
            captured.update(a=np.array(a), b=np.array(b), c=np.array(c), d=np.array(d),
                            water_mask=np.array(water_mask), edge_mask=np.array(edge_mask), out=np.array(res))
        return res

    utilities.solve_tridiagonal = spy
    try:
        isoneutral.isoneutral_diffusion(state, vs.temp, True)
    finally:
        utilities.solve_tridiagonal = orig
    for k, v in captured.items():
        out["tdma__" + k] = v
    out["dT__temp"], out["dT__dtemp_iso"], out["dT__P_diss_iso"] = map(
        np.array, (vs.temp, vs.dtemp_iso, vs.P_diss_iso))

    isoneutral.isoneutral_diffusion(state, vs.salt, False)
    out["dS__salt"], out["dS__dsalt_iso"], out["dS__P_diss_iso"] = map(
        np.array, (vs.salt, vs.dsalt_iso, vs.P_diss_iso))

    isoneutral.isoneutral_skew_diffusion(state, vs.temp, True)
    out["kT__temp"], out["kT__dtemp_iso"], out["kT__P_diss_skew"] = map(
        np.array, (vs.temp, vs.dtemp_iso, vs.P_diss_skew))
    isoneutral.isoneutral_skew_diffusion(state, vs.salt, False)
    out["kS__salt"], out["kS__dsalt_iso"], out["kS__P_diss_skew"] = map(
        np.array, (vs.salt, vs.dsalt_iso, vs.P_diss_skew))


def random_case(name, seed, **extra):
    """test/pyom_consistency/isoneutral_test.py:9-19 settings on a small grid (seed as test/conftest.py:31-35)."""
    from veros.pyom_compat import get_random_state

    settings = dict(
        dt_tracer=3600, dt_mom=3600, enable_neutral_diffusion=True, enable_skew_diffusion=True,
        enable_TEM_friction=True, K_iso_steep=1, enable_streamfunction=False,
    )
    settings.update(extra)
    np.random.seed(seed)
    state = get_random_state(extra_settings=settings)
    out = {}
    snapshot_inputs(state, out)
    run_hot_path(state, out)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def acc_case(name="acc_30x42x15", nsteps=40):
    """BASELINE config 1: veros/setups/acc spun up `nsteps`, state captured at thermodynamics.py:430."""
    from veros.setups.acc import ACCSetup
    from veros.core import isoneutral

    sim = ACCSetup()
    sim.setup()
    state = sim.state
    for _ in range(nsteps):
        sim.step(state)

    out = {}
    orig_pre = isoneutral.isoneutral_diffusion_pre

    class Captured(Exception):
        pass

    def spy_pre(st):
        snapshot_inputs(st, out)
        raise Captured

    isoneutral.isoneutral_diffusion_pre = spy_pre
    try:
        sim.step(state)
    except Captured:
        pass
    finally:
        isoneutral.isoneutral_diffusion_pre = orig_pre

    # rebuild an unlocked state holding the captured inputs and run the path on it
    vs = state.variables
    with vs.unlock():
        for k in INPUT_VARS:
            if "in__" + k in out:
                setattr(vs, k, out["in__" + k].copy())
        vs.P_diss_skew = vs.P_diss_skew * 0.0  # thermodynamics.py:435
        out["in__P_diss_skew"] = np.array(vs.P_diss_skew)
        run_hot_path(state, out)
    # Outside the write regions Ai_* still hold their allocation zeros in a model
    # run, so the previous-step arrays need not be shipped: tests start from zeros instead.
    for k in ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by"):
        prev = out.pop("in__" + k)
        touched = np.zeros(prev.shape, dtype=bool)
        if k == "Ai_ez":
            touched[1:-2, 2:-2, :, :, 1] = True
            touched[1:-2, 2:-2, 1:, :, 0] = True
        elif k == "Ai_nz":
            touched[2:-2, 1:-2, :, :, 1] = True
            touched[2:-2, 1:-2, 1:, :, 0] = True
        else:
            touched[2:-2, 2:-2, :-1] = True
        assert np.all(prev[~touched] == 0.0), k
        out["zeroinit__" + k] = np.int32(1)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    which = sys.argv[1:] or ["random", "acc"]
    if "random" in which:
        random_case("rand_eos1_10x8x7", 17, nx=10, ny=8, nz=7, eq_of_state_type=1)
        random_case("rand_eos2_6x5x4", 18, nx=6, ny=5, nz=4, eq_of_state_type=2)
        random_case("rand_eos3_8x9x6_cyclic", 19, nx=8, ny=9, nz=6, eq_of_state_type=3, enable_cyclic_x=True)
        random_case("rand_eos4_5x6x5", 20, nx=5, ny=6, nz=5, eq_of_state_type=4)
        random_case("rand_eos5_9x7x8", 21, nx=9, ny=7, nz=8, eq_of_state_type=5)
        random_case("rand_eos1_noenergy_7x6x9", 22, nx=7, ny=6, nz=9, eq_of_state_type=1,
                    enable_conserve_energy=False)
    if "acc" in which:
        acc_case()
