"""The reference arm of bench.py: the UNMODIFIED reference (team-ocean/veros, NumPy backend) timed on the
host cores of the box, through its own public API and stock code path:

    vs.update(isoneutral.isoneutral_diffusion_pre(state))
    isoneutral.isoneutral_diffusion(state, vs.temp, True)
    isoneutral.isoneutral_diffusion(state, vs.salt, False)        (veros/core/thermodynamics.py:430-432)

The reference is imported from ``baseline/_ref`` (installed once in the build container with
``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>``,
git-ignored, travels to the GPU box with the snapshot).  None of this repository's kernels, oracle or engine
is on this path; ``veros_b200.synthetic`` only provides the synthetic input state (plain NumPy arrays of the
same workload the GPU arm runs).

JAX is not installed in this image, so the reference's JAX-CPU path cannot be timed.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

STATE_VARS = (
    "temp", "salt", "K_iso", "K_gm", "maskT", "maskU", "maskV", "maskW", "kbot",
    "dxt", "dxu", "dyt", "dyu", "cost", "cosu", "dzt", "dzw", "zt",
    "Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33",
    "dtemp_iso", "dsalt_iso", "P_diss_iso", "P_diss_skew", "int_drhodT", "int_drhodS",
)
SETTINGS = ("eq_of_state_type", "enable_conserve_energy", "enable_cyclic_x", "K_iso_steep", "iso_slopec",
            "iso_dslope", "dt_tracer", "grav", "rho_0")


def available():
    return os.path.isdir(os.path.join(REF, "veros"))


def _import_reference():
    os.environ.setdefault("VEROS_BACKEND", "numpy")
    os.environ.setdefault("VEROS_LOGLEVEL", "error")
    os.environ.setdefault("VEROS_DISKLESS_MODE", "1")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import veros  # noqa: F401
    from veros import runtime_settings as rs

    assert rs.backend == "numpy", rs.backend
    return veros


def reference_state(st):
    """A veros state (reference's own VerosState) holding the synthetic state dict `st`."""
    _import_reference()
    import numpy as np
    from veros.state import get_default_state

    state = get_default_state()
    with state.settings.unlock():
        state.settings.update(nx=int(st["nx"]), ny=int(st["ny"]), nz=int(st["nz"]), enable_neutral_diffusion=True,
                              enable_skew_diffusion=True, **{k: st[k] for k in SETTINGS if k in st})
    state.initialize_variables()
    vs = state.variables
    vs.__locked__ = False
    for name in STATE_VARS:
        if name in st:
            ref = getattr(vs, name)
            setattr(vs, name, np.ascontiguousarray(st[name]).astype(ref.dtype, copy=True))
    vs.tau, vs.taup1 = int(st["tau"]), int(st["taup1"])
    vs.taum1 = 3 - vs.tau - vs.taup1
    return state


def step(state):
    from veros.core import isoneutral

    vs = state.variables
    vs.update(isoneutral.isoneutral_diffusion_pre(state))
    isoneutral.isoneutral_diffusion(state, vs.temp, True)
    isoneutral.isoneutral_diffusion(state, vs.salt, False)


def time_steps(st, steps, warmup):
    """(seconds per step: best, mean) of `steps` timed steps after `warmup` untimed ones."""
    state = reference_state(st)
    for _ in range(warmup):
        step(state)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step(state)
        times.append(time.perf_counter() - t0)
    return min(times), sum(times) / len(times)


def outputs(st):
    """One step of the reference on `st`; returns the path's outputs (used by tests/ in the build container to
    check that this runner feeds the reference the same state the golden fixtures were made from)."""
    import numpy as np

    state = reference_state(st)
    step(state)
    vs = state.variables
    names = ("temp", "salt", "dtemp_iso", "dsalt_iso", "P_diss_iso", "Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33")
    return {k: np.array(getattr(vs, k)) for k in names}
