"""Shared helpers for the parity tests: fixture loading and the parity metrics of SURVEY.md 8(c)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

AI = ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by")
KS = ("K_11", "K_22", "K_33")


def golden_names():
    names = (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return sorted(n for n in names if not n.startswith(("vmix_", "fric_", "sf_", "eke_", "adv_")))


def load_golden(name):
    """Returns (state dict of inputs keyed by reference variable/setting names, dict of stage outputs)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    st, stages = {}, {}
    for key in z.files:
        group, var = key.split("__", 1)
        if group in ("in", "set"):
            st[var] = z[key] if z[key].ndim else z[key].item()
        elif group == "zeroinit":
            pass
        else:
            stages.setdefault(group, {})[var] = z[key]
    for k in AI:  # ACC fixture: Ai_* start from their allocation zeros (see make_golden.py)
        if k not in st:
            st[k] = np.zeros(st["K_iso"].shape + (2, 2))
    return st, stages


def vmix_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("vmix_") and f.endswith(".npz"))


def io_golden_names(prefix):
    """Fixtures of the neighbouring kernels (make_golden_next.py): fric_*, sf_*, eke_*."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith(prefix) and f.endswith(".npz"))


def load_io_golden(name):
    """(inputs and settings keyed by reference names, outputs) of an in__/set__/out__ fixture."""
    return load_vmix_golden(name)


def load_vmix_golden(name):
    """(inputs keyed by reference names, outputs) of a vertmix_tempsalt fixture (make_golden_vmix.py)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    st, out = {}, {}
    for key in z.files:
        group, var = key.split("__", 1)
        val = z[key] if z[key].ndim else z[key].item()
        (out if group == "out" else st)[var] = val
    return st, out


def copy_state(st):
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in st.items()}


def interior(x):
    return x[2:-2, 2:-2]


def norm_err(x, ref):
    """max|x-ref| / max|ref| over the whole array (the reference's _normalize idea, test_base.py:8-17)."""
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max()
    err = np.abs(x - ref).max() if x.size else 0.0
    if scale == 0.0:
        return err
    return err / scale


def tendency_err(dx, dref, dt, tr_ref):
    """dt*max|d(dtracer)| / max|tr|: the tendency error expressed as the tracer change it causes
    (the raw normalised error of (new-old)/dt is cancellation-limited)."""
    return dt * np.abs(np.asarray(dx) - np.asarray(dref)).max() / np.abs(tr_ref).max()
