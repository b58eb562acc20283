"""How the ONE fused step op reaches an unmodified Veros: a stand-in for the module object ``isoneutral`` that
``veros/core/thermodynamics.py`` calls at :430-432

    vs.update(isoneutral.isoneutral_diffusion_pre(state))
    vs.update(isoneutral.isoneutral_diffusion(state, tr=vs.temp, istemp=True))
    vs.update(isoneutral.isoneutral_diffusion(state, tr=vs.salt, istemp=False))

``thermodynamics.isoneutral`` is a module global of thermodynamics.py (``from veros.core import ... isoneutral``), so
rebinding it changes what these three lines call and nothing else in Veros.  The facade's
``isoneutral_diffusion_pre`` runs the whole step and returns ALL twelve arrays the three calls produce (``vs.update``
accepts any subset of variables, veros/state.py:70-90); its ``isoneutral_diffusion`` then has nothing left to do and
returns None, which ``vs.update`` ignores -- exactly what the reference's own ``@veros_routine`` versions return.
Every other attribute (``isoneutral_skew_diffusion``, ``isoneutral_friction``, ``check_isoneutral_slope_crit`` ...) is
looked up on the real package, so later rebinding of those by ``jax_glue.install`` is seen through the facade.

This module has no JAX or CUDA dependency: ``jax_glue.install(fused=True)`` supplies the fused-step kernel; the test
suite drives the same facade with the reference's NumPy functions composed into one call to check the mechanism
against the real thermodynamics.py (tests/test_facade_reference.py).
"""

STEP_OUTPUTS = ("temp", "salt", "dtemp_iso", "dsalt_iso", "P_diss_iso",
                "Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K_11", "K_22", "K_33")


class FusedIsoneutralFacade:
    def __init__(self, iso_package, fused_step):
        """`iso_package`: the real ``veros.core.isoneutral``; `fused_step(state)` -> KernelOutput with STEP_OUTPUTS
        (without P_diss_iso when enable_conserve_energy is off)."""
        self.__dict__["_pkg"] = iso_package
        self.__dict__["_fused_step"] = fused_step

    def __getattr__(self, name):
        return getattr(self.__dict__["_pkg"], name)

    def isoneutral_diffusion_pre(self, state):
        return self.__dict__["_fused_step"](state)

    def isoneutral_diffusion(self, state, tr, istemp):
        """thermodynamics.py:431-432: already done by the fused step of :430."""
        return None


def install_facade(thermodynamics_module, iso_package, fused_step):
    """Rebind ``thermodynamics.isoneutral``; returns the facade (``uninstall_facade`` restores the package)."""
    facade = FusedIsoneutralFacade(iso_package, fused_step)
    thermodynamics_module.isoneutral = facade
    return facade


def uninstall_facade(thermodynamics_module, iso_package):
    thermodynamics_module.isoneutral = iso_package
