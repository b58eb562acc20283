"""Pins the CPU oracle (oracle/iso_oracle.c) against golden vectors produced by the reference's own
NumPy implementation (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import helpers
from helpers import AI, KS, copy_state, golden_names, load_golden, norm_err
from oracle import oracle

NAMES = golden_names()

# isoneutral_diffusion_pre cannot be bit-exact across libm implementations: NumPy evaluates tanh with
# its own SIMD routine, the oracle with glibc's (they differ by 1 ulp on ~25 % of arguments).
PRE_TOL = 2e-15


def test_fixtures_present():
    assert len(NAMES) >= 7
    assert "acc_30x42x15" in NAMES


@pytest.mark.parametrize("name", NAMES)
def test_pre_matches_reference(name):
    st, stages = load_golden(name)
    oracle.isoneutral_diffusion_pre(st)
    for k in AI + KS:
        assert norm_err(st[k], stages["pre"][k]) <= PRE_TOL, k
    # untouched elements keep their previous values bit-for-bit, K_33 top level is zeroed everywhere
    assert np.all(st["K_33"][:, :, -1] == 0.0)


def _state_after_pre(name):
    st, stages = load_golden(name)
    for k in AI + KS:  # feed the reference's own pre outputs: identical inputs for the diffusion op
        st[k] = stages["pre"][k].copy()
    return st, stages


@pytest.mark.parametrize("name", NAMES)
def test_diffusion_bitexact_on_reference_inputs(name):
    st, stages = _state_after_pre(name)
    energy = bool(st["enable_conserve_energy"])
    oracle.isoneutral_diffusion(st, "temp")
    for k in ("flux_east", "flux_north", "flux_top"):
        assert np.array_equal(st[k], stages["fluxT"][k]), k
    assert np.array_equal(st["temp"], stages["dT"]["temp"])
    assert np.array_equal(st["dtemp_iso"], stages["dT"]["dtemp_iso"])
    if energy:
        assert np.array_equal(st["P_diss_iso"], stages["dT"]["P_diss_iso"])
    oracle.isoneutral_diffusion(st, "salt")
    assert np.array_equal(st["salt"], stages["dS"]["salt"])
    assert np.array_equal(st["dsalt_iso"], stages["dS"]["dsalt_iso"])
    if energy:
        assert np.array_equal(st["P_diss_iso"], stages["dS"]["P_diss_iso"])
    oracle.isoneutral_skew_diffusion(st, "temp")
    assert np.array_equal(st["temp"], stages["kT"]["temp"])
    assert np.array_equal(st["dtemp_iso"], stages["kT"]["dtemp_iso"])
    oracle.isoneutral_skew_diffusion(st, "salt")
    assert np.array_equal(st["salt"], stages["kS"]["salt"])
    assert np.array_equal(st["dsalt_iso"], stages["kS"]["dsalt_iso"])
    if energy:
        assert np.array_equal(st["P_diss_skew"], stages["kS"]["P_diss_skew"])


@pytest.mark.parametrize("name", NAMES)
def test_solve_tridiagonal_bitexact_vs_dgtsv(name):
    _, stages = load_golden(name)
    t = stages["tdma"]
    out = oracle.solve_tridiagonal(t["a"], t["b"], t["c"], t["d"], t["water_mask"], t["edge_mask"], mode=0)
    assert np.array_equal(out, t["out"])
    # the Thomas recurrence of the reference's Cython/CUDA kernels agrees to rounding
    out1 = oracle.solve_tridiagonal(t["a"], t["b"], t["c"], t["d"], t["water_mask"], t["edge_mask"], mode=1)
    assert norm_err(out1, t["out"]) < 1e-14


@pytest.mark.parametrize("name", NAMES)
def test_full_step_chained(name):
    """pre -> T -> S with the oracle's own pre outputs (tanh differs by ulps from NumPy's)."""
    st, stages = load_golden(name)
    ref_in = copy_state(st)
    oracle.isoneutral_step(st)
    dt = float(st["dt_tracer"])
    assert norm_err(st["temp"], stages["dT"]["temp"]) < 1e-14
    assert norm_err(st["salt"], stages["dS"]["salt"]) < 1e-14
    for tr, d, stage in (("temp", "dtemp_iso", "dT"), ("salt", "dsalt_iso", "dS")):
        err = dt * np.abs(st[d] - stages[stage][d]).max() / np.abs(ref_in[tr]).max()
        assert err < 1e-13, (d, err)


def test_solve_implicit_edge_overrides_and_empty():
    rng = np.random.default_rng(3)
    X, Y, nz = 5, 4, 9
    a, c = rng.uniform(-0.4, 0, (2, X, Y, nz))
    b = 1.0 + rng.uniform(0.5, 1.0, (X, Y, nz))
    d, b_edge, d_edge = rng.standard_normal((3, X, Y, nz))
    b_edge = 1.5 + np.abs(b_edge)
    kbot = rng.integers(0, nz + 1, (X, Y))
    ks = kbot - 1
    kk = np.arange(nz)[None, None, :]
    land = (ks >= 0)[..., None]
    water, edge = land & (kk >= ks[..., None]), land & (kk == ks[..., None])
    out = oracle.solve_implicit(a, b, c, d, water, edge, b_edge=b_edge, d_edge=d_edge)
    # dense check column by column
    for i in range(X):
        for j in range(Y):
            w = water[i, j]
            n = int(w.sum())
            assert np.all(out[i, j][~w] == 0.0)
            if n == 0:
                continue
            bb = np.where(edge[i, j], b_edge[i, j], b[i, j])[w]
            dd = np.where(edge[i, j], d_edge[i, j], d[i, j])[w]
            A = np.diag(bb) + np.diag(a[i, j][w][1:], -1) + np.diag(c[i, j][w][:-1], 1)
            np.testing.assert_allclose(out[i, j][w], np.linalg.solve(A, dd), rtol=1e-12, atol=1e-13)
    assert oracle.solve_implicit(*(np.zeros((0, 3, nz)),) * 4, np.zeros((0, 3, nz), bool), np.zeros((0, 3, nz), bool)).shape == (0, 3, nz)


def test_random_systems_bitexact_vs_scipy_dgtsv():
    """test/pyom_consistency/tridiag_test.py:8-37 inputs (70x60x50 randn systems, random kbot): with
    dgtsv's partial pivoting mirrored the oracle reproduces SciPy's LAPACK bit-for-bit even on systems
    that are not diagonally dominant (call sequence of veros/core/operators.py:60-77)."""
    from scipy.linalg import lapack

    nx, ny, nz = 70, 60, 50
    a, b, c, d = np.random.randn(4, nx, ny, nz)
    kbot = np.random.randint(0, nz, size=(nx, ny))
    ks = kbot - 1
    kk = np.arange(nz)[None, None, :]
    land = (ks >= 0)[..., None]
    water, edge = land & (kk >= ks[..., None]), land & (kk == ks[..., None])
    aa, cc = a.copy(), c.copy()
    aa[edge] = 0
    cc[..., -1] = 0
    ref = np.zeros_like(a)
    ref[water] = lapack.dgtsv(aa[water][1:], b[water], cc[water][:-1], d[water])[3]
    out = oracle.solve_tridiagonal(a, b, c, d, water, edge, mode=0)
    assert np.array_equal(out, ref)


# ------------------------------------------------------------------ vertmix_tempsalt
@pytest.mark.parametrize("name", helpers.vmix_golden_names())
def test_oracle_vertmix_tempsalt_bitexact(name):
    """veros/core/thermodynamics.py:248-300 incl. enforce_boundaries: bit for bit, also where dgtsv
    interchanges rows (random kappaH is not positive)."""
    st, out = helpers.load_vmix_golden(name)
    got = oracle.vertmix_tempsalt(helpers.copy_state(st))
    for k in ("temp", "salt", "dtemp_vmix", "dsalt_vmix"):
        assert np.array_equal(got[k], out[k]), k


# ------------------------------------------------------------------ neighbours of the path
@pytest.mark.parametrize("name", helpers.io_golden_names("fric_"))
def test_oracle_implicit_vert_friction_bitexact(name):
    """veros/core/friction.py:92-205 (two solve_implicit calls with b_edge, tendencies, dissipation through
    ugrid_to_tgrid / vgrid_to_tgrid): bit for bit against the reference's NumPy backend."""
    st, out = helpers.load_io_golden(name)
    got = oracle.implicit_vert_friction(helpers.copy_state(st))
    for k in ("u", "v", "du_mix", "dv_mix", "K_diss_v"):
        assert np.array_equal(got[k], out[k]), k


@pytest.mark.parametrize("name", helpers.io_golden_names("sf_"))
def test_oracle_diag_streamfunction_bitexact(name):
    st, out = helpers.load_io_golden(name)
    got = oracle.isoneutral_diag_streamfunction(helpers.copy_state(st))
    for k in ("B1_gm", "B2_gm"):
        assert np.array_equal(got[k], out[k]), k


@pytest.mark.parametrize("name", helpers.io_golden_names("eke_"))
def test_oracle_set_eke_diffusivities_bitexact(name):
    """veros/core/eke.py:34-85 incl. NumPy's pairwise order of the column sum (nz = 9: 8-accumulator block + tail,
    nz = 140: recursive halving) and the enable_eke = False branch."""
    st, out = helpers.load_io_golden(name)
    got = oracle.set_eke_diffusivities(helpers.copy_state(st))
    assert set(got) == set(out)
    for k in out:
        assert np.array_equal(got[k], out[k]), k


@pytest.mark.parametrize("name", helpers.io_golden_names("adv_"))
def test_oracle_advect_tempsalt_bitexact(name):
    """advect_tracer with the superbee / 2nd-order fluxes (thermodynamics.py:10-40, advection.py:8-115),
    advect_temperature / advect_salinity and the Adams-Bashforth step (:223-245): bit for bit."""
    st, out = helpers.load_io_golden(name)
    assert np.array_equal(oracle.advect_tracer(helpers.copy_state(st), "temp"), out["dtr_temp"])
    got = oracle.advect_tempsalt(helpers.copy_state(st))
    for k in ("temp", "salt", "dtemp", "dsalt"):
        assert np.array_equal(got[k], out[k]), k
