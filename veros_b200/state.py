"""Device-resident state for the isoneutral path.

Mirrors the two containers the reference's kernels read (veros/state.py: ``state.variables`` and
``state.settings``) with exactly the variable / setting names of veros/variables.py and
veros/settings.py that the hot path touches, so host code reads like the reference's:

    vs = state.variables
    vs.update(isoneutral.isoneutral_diffusion_pre(state))

Arrays are torch CUDA tensors used purely as device buffers (dtype / shape / data_ptr); no torch
arithmetic is ever applied to them.
"""
import weakref
from collections import namedtuple
from types import SimpleNamespace

import numpy as np
import torch

F64_3D = ("K_iso", "K_gm", "K_11", "K_22", "K_33", "dtemp_iso", "dsalt_iso", "P_diss_iso", "P_diss_skew",
          "kappaH", "dtemp_vmix", "dsalt_vmix",
          # neighbours of the path (friction.py, eke.py, isoneutral_diag_streamfunction)
          "kappaM", "du_mix", "dv_mix", "K_diss_v", "B1_gm", "B2_gm", "L_rhines", "eke_len", "sqrteke")
F64_2D = ("forc_temp_surface", "forc_salt_surface", "area_v", "area_t", "coriolis_t", "beta", "L_rossby")
F64_4D = ("temp", "salt", "int_drhodT", "int_drhodS", "u", "v", "Nsqr", "eke", "w", "dtemp", "dsalt")
F64_5D = ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by")
MASKS = ("maskT", "maskU", "maskV", "maskW")
METRICS_X = ("dxt", "dxu")
METRICS_Y = ("dyt", "dyu", "cost", "cosu")
METRICS_Z = ("dzt", "dzw", "zt")
SETTINGS = ("eq_of_state_type", "enable_conserve_energy", "K_iso_steep", "iso_slopec", "iso_dslope",
            "dt_tracer", "grav", "rho_0")
EXTRA_SETTINGS = ("AB_eps", "enable_superbee_advection", "dt_mom", "enable_eke", "enable_eke_isopycnal_diffusion", "pi", "eke_lmin", "eke_cross", "eke_crhin",
                  "eke_k_max", "eke_c_k", "K_gm_0", "K_iso_0")
OPTIONAL = ("K_gm", "P_diss_skew", "int_drhodT", "int_drhodS", "P_diss_iso",
            # vertmix_tempsalt (veros_b200/thermodynamics.py)
            "kappaH", "dtemp_vmix", "dsalt_vmix", "forc_temp_surface", "forc_salt_surface",
            # friction.py, eke.py, isoneutral_diag_streamfunction
            "kappaM", "du_mix", "dv_mix", "K_diss_v", "B1_gm", "B2_gm", "L_rhines", "eke_len", "sqrteke", "area_v", "area_t",
            "coriolis_t", "beta", "L_rossby", "u", "v", "Nsqr", "eke", "w", "dtemp", "dsalt")


def KernelOutput(**kwargs):
    """Same factory as veros.state.KernelOutput (veros/state.py:16-20)."""
    return namedtuple("KernelOutput", list(kwargs.keys()))(*kwargs.values())


class Variables(SimpleNamespace):
    def update(self, out):
        """vs.update(KernelOutput) as in veros/state.py."""
        for k, v in out._asdict().items():
            setattr(self, k, v)


class IsoState:
    """``state.variables`` / ``state.settings`` for the isoneutral path, on one GPU."""

    def __init__(self, variables, settings, device):
        self.variables = variables
        self.settings = settings
        self.device = device
        self._workspace = None
        self._dummy = torch.zeros(8, dtype=torch.float64, device=device)
        self.ring_flags = 0  # VEROS_B200_FLAG_NO_WEST_RING / _NO_EAST_RING for x-sub-slab views
        self.tuning_flags = 0  # VEROS_B200_FLAG_PRE_SINGLE / _PRE_SPLIT ...: kernel-variant knobs (tests, tuning)
        self._views = weakref.WeakSet()  # sub-slab views: advance_time() keeps their host-side time levels in step

    # ---- construction --------------------------------------------------------------------------
    @classmethod
    def from_numpy(cls, st, device="cuda", strict=True):
        """`st`: dict keyed by reference variable / setting names (see tests/helpers.py).
        strict=False accepts a state that only carries the variables of one of the neighbouring
        kernels (e.g. vertmix_tempsalt); the op wrappers check for what they need."""
        device = torch.device(device)
        ref = next(st[n] for n in F64_4D + F64_3D if n in st)  # any cell-shaped array fixes the grid
        N, M, nz = np.shape(ref)[:3]
        vs = Variables()

        def put(name, arr, dtype):
            arr = np.ascontiguousarray(arr, dtype=dtype)
            setattr(vs, name, torch.from_numpy(arr).to(device))

        for name in F64_3D + F64_2D + F64_4D + F64_5D + METRICS_X + METRICS_Y + METRICS_Z:
            if name in st:
                put(name, st[name], np.float64)
            elif strict and name not in OPTIONAL:
                raise KeyError(name)
        for name in MASKS:
            if name in st or strict:
                put(name, np.asarray(st[name]).astype(np.uint8), np.uint8)
        if "kbot" in st or strict:
            put("kbot", st["kbot"], np.int32)
        for name in ("tau", "taup1", "taum1"):
            if name in st or (strict and name != "taum1"):
                put(name, np.array([int(st[name])]), np.int32)
                setattr(vs, name + "_host", int(st[name]))  # host copy for plumbing (halo packing)
        settings = SimpleNamespace(**{k: st[k] for k in SETTINGS if strict or k in st})
        for k in EXTRA_SETTINGS:  # settings of the neighbouring kernels, kept when present
            if k in st:
                setattr(settings, k, st[k].item() if isinstance(st[k], np.ndarray) else st[k])
        if hasattr(settings, "eq_of_state_type"):
            settings.eq_of_state_type = int(settings.eq_of_state_type)
        if hasattr(settings, "enable_conserve_energy"):
            settings.enable_conserve_energy = bool(settings.enable_conserve_energy)
        for k in SETTINGS[2:]:
            if hasattr(settings, k):
                setattr(settings, k, float(getattr(settings, k)))
        settings.enable_cyclic_x = bool(st.get("enable_cyclic_x", False))
        settings.nx, settings.ny, settings.nz = N - 4, M - 4, nz
        obj = cls(vs, settings, device)
        obj.validate(strict)
        return obj

    def validate(self, strict=True):
        vs, st = self.variables, self.settings
        N, M, nz = st.nx + 4, st.ny + 4, st.nz
        shapes = {**{n: (N, M, nz) for n in F64_3D + MASKS}, **{n: (N, M, nz, 3) for n in F64_4D},
                  **{n: (N, M) for n in F64_2D},
                  **{n: (N, M, nz, 2, 2) for n in F64_5D}, **{n: (N,) for n in METRICS_X},
                  **{n: (M,) for n in METRICS_Y}, **{n: (nz,) for n in METRICS_Z}, "kbot": (N, M),
                  "tau": (1,), "taup1": (1,)}
        for name, shape in shapes.items():
            t = getattr(vs, name, None)
            if t is None:
                if name in OPTIONAL or not strict:
                    continue
                raise ValueError(f"state is missing variable {name}")
            if tuple(t.shape) != shape:
                raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {shape}")
            want = torch.uint8 if name in MASKS else torch.int32 if name in ("kbot", "tau", "taup1") else torch.float64
            if t.dtype != want:
                raise TypeError(f"{name} has dtype {t.dtype}, expected {want}")
            if not t.is_contiguous():
                raise ValueError(f"{name} must be C-contiguous")
        if not strict:
            return
        if st.enable_conserve_energy:
            for name in ("int_drhodT", "int_drhodS", "P_diss_iso"):
                if getattr(vs, name, None) is None:
                    raise ValueError(f"enable_conserve_energy needs variable {name}")
        if not 1 <= st.eq_of_state_type <= 5:
            raise ValueError("unknown equation of state")

    # ---- time stepping ---------------------------------------------------------------------------
    def advance_time(self):
        """Rotate the time-level indices like the end of a model step (veros/veros.py: taum1, tau, taup1 =
        tau, taup1, taum1): the device scalars the kernels read and the host copies the plumbing uses (halo
        exchange level) change TOGETHER.  Sub-slab views share both with their parent."""
        vs = self.variables
        tau, taup1 = int(vs.tau_host), int(vs.taup1_host)
        taum1 = 3 - tau - taup1
        vs.tau_host, vs.taup1_host = taup1, taum1
        vs.tau.fill_(taup1)
        vs.taup1.fill_(taum1)
        for sub in list(self._views):
            sub.variables.tau_host, sub.variables.taup1_host = taup1, taum1

    # ---- helpers --------------------------------------------------------------------------------
    def to_numpy(self, names=None):
        vs = self.variables
        names = names or [n for n in vars(vs) if not n.endswith("_host")]
        out = {}
        for n in names:
            t = getattr(vs, n)
            out[n] = t.cpu().numpy()
        return out

    def subslab(self, i0, i1):
        """View of the x-planes [i0, i1) (ghost planes included: the view's interior is [i0+2, i1-2)).
        x is the slowest axis, so every sliced array stays contiguous and the ops run on it unchanged.
        Interior sub-slabs of a wider slab get the ring flags so that P_diss is accumulated exactly once
        (include/veros_b200.h, VEROS_B200_FLAG_NO_WEST_RING)."""
        from . import _lib

        N = self.settings.nx + 4
        if not (0 <= i0 and i1 <= N and i1 - i0 >= 5):
            raise ValueError(f"bad sub-slab [{i0}, {i1}) of {N} planes")
        vs = Variables()
        for name, t in vars(self.variables).items():
            if name in ("tau", "taup1", "tau_host", "taup1_host") or name in METRICS_Y or name in METRICS_Z:
                setattr(vs, name, t)
            else:
                setattr(vs, name, t[i0:i1])
        settings = SimpleNamespace(**vars(self.settings))
        settings.nx = (i1 - i0) - 4
        sub = IsoState(vs, settings, self.device)
        sub._dummy = self._dummy
        sub.tuning_flags = self.tuning_flags
        sub.ring_flags = (0 if i0 == 0 else _lib.FLAG_NO_WEST_RING) | (0 if i1 == N else _lib.FLAG_NO_EAST_RING)
        sub._parent = self
        self._views.add(sub)
        return sub

    def workspace(self, nbytes):
        parent = getattr(self, "_parent", None)
        if parent is not None:  # sub-slabs run one after the other on a stream: share the parent's scratch
            return parent.workspace(nbytes)
        n = max(8, (int(nbytes) + 7) // 8)
        if self._workspace is None or self._workspace.numel() < n:
            self._workspace = torch.empty(n, dtype=torch.float64, device=self.device)
        return self._workspace

    def dummy(self):
        return self._dummy
