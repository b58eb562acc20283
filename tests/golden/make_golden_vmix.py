#!/usr/bin/env python
"""Golden vectors for ``vertmix_tempsalt`` (veros/core/thermodynamics.py:248-300), the step right
after the isoneutral path, from the REFERENCE's NumPy implementation.

Runs only in the build container (imports veros from /root/reference); the ``vmix_*.npz`` files it
writes are committed.

    python tests/golden/make_golden_vmix.py

Each fixture holds the inputs of the kernel (``in__*``: temp, salt, kappaH, forc_*_surface, kbot, dzt,
dzw, taup1; ``set__*``: nx, ny, nz, dt_tracer, enable_cyclic_x) and its four outputs (``out__temp``,
``out__salt``, ``out__dtemp_vmix``, ``out__dsalt_vmix``), the latter including the cyclic / closed
boundary treatment of ``enforce_boundaries`` (:290-297).  The random states have ``kappaH = randn``,
i.e. matrices that are not diagonally dominant, so ``dgtsv`` interchanges rows in many columns.
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
INPUTS = ["temp", "salt", "kappaH", "forc_temp_surface", "forc_salt_surface", "kbot", "dzt", "dzw"]
SETTINGS = ["nx", "ny", "nz", "dt_tracer", "enable_cyclic_x"]


def snapshot(state, out):
    vs, st = state.variables, state.settings
    for name in INPUTS:
        out["in__" + name] = np.array(getattr(vs, name))
    out["in__taup1"] = np.int32(vs.taup1)
    for name in SETTINGS:
        out["set__" + name] = np.asarray(getattr(st, name))


def run(state, out):
    from veros.core import thermodynamics

    vs = state.variables
    vs.update(thermodynamics.vertmix_tempsalt(state))
    for name in ("temp", "salt", "dtemp_vmix", "dsalt_vmix"):
        out["out__" + name] = np.array(getattr(vs, name))


def random_case(name, seed, **extra):
    from veros.pyom_compat import get_random_state

    settings = dict(dt_tracer=3600, dt_mom=3600, enable_streamfunction=False)
    settings.update(extra)
    np.random.seed(seed)
    state = get_random_state(extra_settings=settings)
    out = {}
    snapshot(state, out)
    run(state, out)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e3:.0f} kB")


def acc_case(name="vmix_acc_30x42x15", nsteps=40):
    """veros/setups/acc spun up `nsteps`; inputs captured at the vertmix_tempsalt call of the next step."""
    from veros.setups.acc import ACCSetup
    from veros.core import thermodynamics

    sim = ACCSetup()
    sim.setup()
    state = sim.state
    for _ in range(nsteps):
        sim.step(state)
    out = {}
    orig = thermodynamics.vertmix_tempsalt

    def spy(st):
        if not out:
            snapshot(st, out)
            res = orig(st)
            for name in ("temp", "salt", "dtemp_vmix", "dsalt_vmix"):
                out["out__" + name] = np.array(getattr(res, name))
            return res
        return orig(st)

    thermodynamics.vertmix_tempsalt = spy
    try:
        sim.step(state)
    finally:
        thermodynamics.vertmix_tempsalt = orig
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    random_case("vmix_rand_10x8x7", 31, nx=10, ny=8, nz=7)
    random_case("vmix_rand_6x5x12_cyclic", 32, nx=6, ny=5, nz=12, enable_cyclic_x=True)
    random_case("vmix_rand_9x7x3", 33, nx=9, ny=7, nz=3)
    acc_case()
