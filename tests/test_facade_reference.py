"""The facade that routes veros/core/thermodynamics.py:430-432 through ONE fused step (veros_b200/facade.py), checked
against the REAL, unmodified reference model: the ACC setup (BASELINE.json configs[0]) is stepped with and without
the facade installed; with it, `isoneutral_diffusion_pre` is a function that performs all three calls and returns all
twelve arrays (here composed from the reference's own NumPy functions -- the mechanism under test is the rebinding and
the `vs.update` protocol, not the arithmetic), and `isoneutral_diffusion` becomes a no-op.  Every prognostic field
must come out bit-identical.  Needs the reference (baseline/_ref, or /root/reference in the build container)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_path():
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "veros")):
            return p
    return None


@pytest.mark.skipif(_reference_path() is None, reason="needs the reference (baseline/_ref)")
def test_model_steps_through_the_facade_are_bit_identical():
    os.environ.setdefault("VEROS_BACKEND", "numpy")
    os.environ.setdefault("VEROS_LOGLEVEL", "error")
    os.environ.setdefault("VEROS_DISKLESS_MODE", "1")
    sys.dont_write_bytecode = True
    if _reference_path() not in sys.path:
        sys.path.insert(0, _reference_path())
    import veros.core.isoneutral as iso_pkg
    import veros.core.thermodynamics as thermodynamics
    from veros import KernelOutput
    from veros.setups.acc import ACCSetup

    from veros_b200 import facade

    calls = {"fused": 0, "noop": 0}

    def fused_step(state):
        """pre + diffusion(temp) + diffusion(salt) in one call, returning every array they touch."""
        calls["fused"] += 1
        vs = state.variables
        vs.update(iso_pkg.isoneutral_diffusion_pre(state))
        iso_pkg.isoneutral_diffusion(state, tr=vs.temp, istemp=True)
        iso_pkg.isoneutral_diffusion(state, tr=vs.salt, istemp=False)
        return KernelOutput(**{k: getattr(vs, k) for k in facade.STEP_OUTPUTS})

    def run(with_facade, nsteps=4):
        sim = ACCSetup()
        sim.setup()
        if with_facade:
            f = facade.install_facade(thermodynamics, iso_pkg, fused_step)
            orig_noop = f.isoneutral_diffusion

            def counting(state, tr, istemp):
                calls["noop"] += 1
                return orig_noop(state, tr, istemp)

            f.__dict__["isoneutral_diffusion"] = counting
        try:
            for _ in range(nsteps):
                sim.step(sim.state)
        finally:
            facade.uninstall_facade(thermodynamics, iso_pkg)
        vs = sim.state.variables
        return {k: np.array(getattr(vs, k)) for k in ("temp", "salt", "u", "v", "K_33", "Ai_ez", "dtemp_iso", "P_diss_iso", "psi")}

    plain = run(False)
    fused = run(True)
    assert calls["fused"] == 4 and calls["noop"] == 8  # one fused call and two no-op calls per model step
    assert thermodynamics.isoneutral is iso_pkg          # uninstalled again
    for k, v in plain.items():
        assert np.array_equal(v, fused[k]), k
    # the facade forwards everything else to the real package
    f = facade.FusedIsoneutralFacade(iso_pkg, fused_step)
    assert f.isoneutral_skew_diffusion is iso_pkg.isoneutral_skew_diffusion
