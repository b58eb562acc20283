"""GPU parity of the neighbours of the isoneutral path through the C ABI:
implicit_vert_friction (fused coefficient assembly + dgtsv replay), isoneutral_diag_streamfunction and
set_eke_diffusivities_kernel -- bit for bit against the reference's golden vectors (tests/golden/fric_*, sf_*,
eke_*.npz, made by make_golden_next.py from the imported reference) and against the CPU oracle on larger seeded
states, including column depths of the benchmark grids."""
import numpy as np
import pytest

import helpers
from helpers import copy_state

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def gpu_state(st, dev):
    from veros_b200.state import IsoState

    return IsoState.from_numpy(st, dev, strict=False)


# ---------------------------------------------------------------------------------- implicit_vert_friction
FRIC_OUT = ("u", "v", "du_mix", "dv_mix", "K_diss_v")


@pytest.mark.parametrize("name", helpers.io_golden_names("fric_"))
def test_implicit_vert_friction_bitexact_vs_reference_golden(name, dev):
    from veros_b200 import friction

    st, out = helpers.load_io_golden(name)
    gs = gpu_state(st, dev)
    res = friction.implicit_vert_friction(gs)
    assert res._fields == FRIC_OUT  # KernelOutput of friction.py:205
    got = gs.to_numpy(list(FRIC_OUT))
    for k in FRIC_OUT:
        assert np.array_equal(got[k], out[k]), k


def friction_state(nx, ny, nz, seed, kappa_sign=False):
    """Random state in the spirit of get_random_state (veros/pyom_compat.py): randn velocities, random kbot with
    islands, masks from kbot; kappaM positive (model-like) or of random sign (interchanges in dgtsv)."""
    from veros_b200 import synthetic

    base = synthetic.random_state(nx, ny, max(nz, 2), seed=seed)
    rng = np.random.default_rng(seed + 100)
    N, M = nx + 4, ny + 4
    st = {k: base[k] for k in ("dxt", "dxu", "tau", "taup1")}
    st["kbot"] = np.minimum(base["kbot"], nz).astype(np.int32)
    _, st["maskU"], st["maskV"], _ = synthetic.masks_from_kbot(st["kbot"], nz, False)
    st["dzt"], st["dzw"] = base["dzt"][:nz].copy(), base["dzw"][:nz].copy()
    st["u"], st["v"] = rng.standard_normal((2, N, M, nz, 3))
    kap = rng.standard_normal((N, M, nz))
    st["kappaM"] = kap if kappa_sign else np.abs(kap) * 1e-2
    st["du_mix"], st["dv_mix"], st["K_diss_v"] = rng.standard_normal((3, N, M, nz))
    st["area_t"] = np.abs(rng.standard_normal((N, M))) + 1.0
    st["area_v"] = np.abs(rng.standard_normal((N, M))) + 1.0
    st["dt_mom"] = 3600.0
    return st


@pytest.mark.parametrize("shape,kappa_sign", [((48, 40, 50), False), ((20, 18, 115), False), ((33, 21, 17), True),
                                              ((6, 5, 1), False), ((150, 7, 2), True)])
def test_implicit_vert_friction_vs_oracle(shape, kappa_sign, dev):
    from oracle import oracle
    from veros_b200 import friction

    st = friction_state(*shape, seed=3, kappa_sign=kappa_sign)
    ref = oracle.implicit_vert_friction(copy_state(st))
    gs = gpu_state(st, dev)
    friction.implicit_vert_friction(gs)
    got = gs.to_numpy(list(FRIC_OUT))
    for k in FRIC_OUT:
        assert np.array_equal(got[k], ref[k]), k
    # the reference's write regions: nothing outside [1:-2, 1:-2] of the velocities and tendencies changes
    for k in ("u", "v", "du_mix", "dv_mix"):
        assert np.array_equal(got[k][-2:], st[k][-2:]) and np.array_equal(got[k][:1], st[k][:1]), k
        assert np.array_equal(got[k][:, -2:], st[k][:, -2:]) and np.array_equal(got[k][:, :1], st[k][:, :1]), k


def test_implicit_vert_friction_errors(dev):
    from veros_b200 import _lib, friction

    st = friction_state(8, 7, 5, seed=1)
    gs = gpu_state(st, dev)
    del gs.variables.kappaM
    with pytest.raises(ValueError):
        friction.implicit_vert_friction(gs)
    with pytest.raises(RuntimeError, match="bad descriptor"):
        _lib.call("veros_b200_implicit_vert_friction_f64", [0] * 23, bytes(8), 0)


# ---------------------------------------------------------------------------------- isoneutral_diag_streamfunction
@pytest.mark.parametrize("name", helpers.io_golden_names("sf_"))
def test_diag_streamfunction_bitexact_vs_reference_golden(name, dev):
    from veros_b200 import isoneutral

    st, out = helpers.load_io_golden(name)
    gs = gpu_state(st, dev)
    res = isoneutral.isoneutral_diag_streamfunction(gs)
    assert res._fields == ("B1_gm", "B2_gm")
    got = gs.to_numpy(["B1_gm", "B2_gm"])
    for k in ("B1_gm", "B2_gm"):
        assert np.array_equal(got[k], out[k]), k


def test_diag_streamfunction_after_the_step_vs_oracle(dev):
    """Chained as in a model step: the fused isoneutral step produces Ai_ez / Ai_nz, the diagnostic consumes them."""
    from oracle import oracle
    from veros_b200 import isoneutral, synthetic
    from veros_b200.state import IsoState

    st = synthetic.make_workload("global_4deg", nx=30, ny=20)
    rng = np.random.default_rng(9)
    st["B1_gm"], st["B2_gm"] = rng.standard_normal((2,) + st["K_gm"].shape)
    gs = IsoState.from_numpy(st, dev)
    isoneutral.isoneutral_step(gs)
    isoneutral.isoneutral_diag_streamfunction(gs)
    got = gs.to_numpy(["Ai_ez", "Ai_nz", "B1_gm", "B2_gm"])
    ref = copy_state(st)
    ref["Ai_ez"], ref["Ai_nz"] = got["Ai_ez"], got["Ai_nz"]  # identical inputs: the diagnostic itself is strict
    oracle.isoneutral_diag_streamfunction(ref)
    for k in ("B1_gm", "B2_gm"):
        assert np.array_equal(got[k], ref[k]), k


# ---------------------------------------------------------------------------------- set_eke_diffusivities_kernel
@pytest.mark.parametrize("name", helpers.io_golden_names("eke_"))
def test_set_eke_diffusivities_bitexact_vs_reference_golden(name, dev):
    from veros_b200 import eke

    st, out = helpers.load_io_golden(name)
    gs = gpu_state(st, dev)
    res = eke.set_eke_diffusivities_kernel(gs)
    assert set(res._fields) == set(out)
    got = gs.to_numpy(list(out))
    for k in out:
        assert np.array_equal(got[k], out[k]), k


@pytest.mark.parametrize("nz", [1, 7, 8, 50, 115, 128, 129, 300])
def test_set_eke_diffusivities_vs_oracle(nz, dev):
    """Column depths around the regime changes of NumPy's pairwise summation (8, 128) and of the benchmark grids."""
    from oracle import oracle
    from veros_b200 import eke

    rng = np.random.default_rng(nz)
    N, M = 11, 9
    st = dict(Nsqr=rng.standard_normal((N, M, nz, 3)) * 1e-5, eke=rng.standard_normal((N, M, nz, 3)) * 1e-2,
              maskW=rng.random((N, M, nz)) < 0.8, dzw=np.abs(rng.standard_normal(nz)) * 50 + 5,
              coriolis_t=rng.standard_normal((N, M)) * 1e-4, beta=np.abs(rng.standard_normal((N, M))) * 2e-11,
              K_gm=np.zeros((N, M, nz)), K_iso=np.zeros((N, M, nz)), tau=1, enable_eke=True,
              enable_eke_isopycnal_diffusion=bool(nz % 2), pi=np.pi, eke_lmin=100.0, eke_cross=2.0, eke_crhin=1.0,
              eke_k_max=1e4, eke_c_k=0.4, K_gm_0=1000.0, K_iso_0=800.0)
    ref = oracle.set_eke_diffusivities(copy_state(st))
    gs = gpu_state(st, dev)
    eke.set_eke_diffusivities_kernel(gs)
    got = gs.to_numpy(list(ref))
    for k in ref:
        assert np.array_equal(got[k], ref[k]), k


# ---------------------------------------------------------------------------------- advect_tempsalt (+ Adams-Bashforth)
ADV_OUT = ("temp", "salt", "dtemp", "dsalt")


@pytest.mark.parametrize("name", helpers.io_golden_names("adv_"))
def test_advect_tempsalt_bitexact_vs_reference_golden(name, dev):
    from veros_b200 import thermodynamics

    st, out = helpers.load_io_golden(name)
    gs = gpu_state(st, dev)
    res = thermodynamics.advect_tempsalt(gs)
    assert res._fields == ADV_OUT
    got = gs.to_numpy(list(ADV_OUT))
    for k in ADV_OUT:
        assert np.array_equal(got[k], out[k]), k
    # tendencies only (advect_temperature / advect_salinity without the time step)
    gs = gpu_state(st, dev)
    thermodynamics.advect_tempsalt(gs, adams_bashforth=False)
    got = gs.to_numpy(list(ADV_OUT))
    assert np.array_equal(got["dtemp"][..., int(st["tau"])], out["dtr_temp"])
    assert np.array_equal(got["temp"], st["temp"]) and np.array_equal(got["dsalt"], out["dsalt"])


@pytest.mark.parametrize("shape,superbee", [((48, 40, 50), True), ((48, 40, 50), False), ((20, 18, 115), True),
                                            ((9, 6, 1), True), ((150, 7, 2), True)])
def test_advect_tempsalt_vs_oracle(shape, superbee, dev):
    from oracle import oracle
    from veros_b200 import synthetic, thermodynamics

    nx, ny, nz = shape
    base = synthetic.random_state(nx, ny, max(nz, 2), seed=5)
    rng = np.random.default_rng(6)
    N, M = nx + 4, ny + 4
    st = {k: base[k] for k in ("dxt", "dyt", "cost", "cosu", "tau", "taup1", "dt_tracer")}
    st["cosu"] = np.cos(np.linspace(-1.2, 1.2, M))  # not all ones: the north flux carries cosu
    st["cost"] = np.cos(np.linspace(-1.15, 1.15, M))
    st["kbot"] = np.minimum(base["kbot"], nz).astype(np.int32)
    st["maskT"], st["maskU"], st["maskV"], st["maskW"] = synthetic.masks_from_kbot(st["kbot"], nz, False)
    st["dzt"] = base["dzt"][:nz].copy()
    st["temp"], st["salt"], st["dtemp"], st["dsalt"] = rng.standard_normal((4, N, M, nz, 3))
    st["u"], st["v"], st["w"] = rng.standard_normal((3, N, M, nz, 3)) * np.array([1.0, 1.0, 1e-3])[:, None, None, None, None]
    st["AB_eps"], st["enable_superbee_advection"] = 0.1, superbee
    ref = oracle.advect_tempsalt(copy_state(st))
    gs = gpu_state(st, dev)
    thermodynamics.advect_tempsalt(gs)
    got = gs.to_numpy(list(ADV_OUT))
    for k in ADV_OUT:
        assert np.array_equal(got[k], ref[k]), k
