"""Host-side logic that needs no GPU: state validation, synthetic generators, error behaviour of the
reference-facing functions, slab decomposition arithmetic."""
import numpy as np
import pytest

from helpers import copy_state

torch = pytest.importorskip("torch")


def small_state(**kw):
    from veros_b200 import synthetic

    return synthetic.random_state(6, 5, 4, **kw)


def test_state_from_numpy_and_validation():
    from veros_b200.state import IsoState

    st = small_state()
    s = IsoState.from_numpy(st, "cpu")
    assert s.settings.nx == 6 and s.settings.nz == 4
    assert s.variables.maskT.dtype == torch.uint8 and s.variables.kbot.dtype == torch.int32
    assert tuple(s.variables.temp.shape) == (10, 9, 4, 3)
    back = s.to_numpy(["temp", "Ai_ez"])
    assert np.array_equal(back["temp"], st["temp"])
    s.variables.K_11 = s.variables.K_11[:, :, :3]
    with pytest.raises(ValueError, match="K_11"):
        s.validate()
    s = IsoState.from_numpy(st, "cpu")
    s.variables.salt = s.variables.salt.float()
    with pytest.raises(TypeError, match="salt"):
        s.validate()
    bad = copy_state(st)
    bad["eq_of_state_type"] = 9
    with pytest.raises(ValueError, match="equation of state"):
        IsoState.from_numpy(bad, "cpu")
    bad = copy_state(st)
    del bad["int_drhodT"]
    with pytest.raises(ValueError, match="int_drhodT"):
        IsoState.from_numpy(bad, "cpu")
    bad["enable_conserve_energy"] = False
    IsoState.from_numpy(bad, "cpu")  # fine without the energy fields


def test_kernel_output_matches_reference_factory():
    from veros_b200.state import KernelOutput, Variables

    out = KernelOutput(K_11=1, K_22=2)
    assert out._fields == ("K_11", "K_22") and out.K_22 == 2
    vs = Variables()
    vs.update(out)
    assert vs.K_11 == 1


def test_ops_refuse_cpu_tensors():
    from veros_b200 import utilities

    x = torch.zeros((2, 3, 5), dtype=torch.float64)
    m = torch.zeros((2, 3, 5), dtype=torch.bool)
    with pytest.raises(RuntimeError, match="no CPU path"):
        utilities.solve_tridiagonal(x, x, x, x, m, m)
    with pytest.raises(ValueError, match="identical shape"):
        utilities.solve_implicit(x, x[:1], x, x, m, m)
    with pytest.raises(TypeError):
        utilities.solve_implicit(*(x.float(),) * 4, m, m)


def test_masks_follow_calc_topo_rules():
    from veros_b200 import synthetic

    st = synthetic.random_state(9, 8, 6, enable_cyclic_x=True)
    kbot, T, U, V, W = (st[k] for k in ("kbot", "maskT", "maskU", "maskV", "maskW"))
    k = np.arange(6)[None, None, :]
    assert np.array_equal(T[2:-2], ((kbot > 0)[..., None] & (kbot[..., None] - 1 <= k))[2:-2])
    assert np.array_equal(U[:-1][2:-2], (T[:-1] & T[1:])[2:-2])
    assert np.array_equal(V[:, :-1], T[:, :-1] & T[:, 1:])
    assert np.array_equal(W[:, :, :-1], T[:, :, :-1] & T[:, :, 1:])
    assert np.array_equal(T[-2:], T[2:4]) and np.array_equal(T[:2], T[-4:-2])  # cyclic ghost columns


def test_analytic_slabs_tile_the_global_state():
    from veros_b200 import decomp, synthetic

    full = synthetic.make_workload("global_4deg")
    nxg = full["nx"]
    for world in (2, 3):
        for rank in range(world):
            x0, x1 = decomp.slab_bounds(nxg, world, rank)
            part = synthetic.make_workload("global_4deg", nx=x1 - x0, x_offset=x0, nx_global=nxg)
            for k in ("temp", "salt", "K_iso", "kbot", "maskT", "int_drhodT"):
                assert np.array_equal(part[k], full[k][x0:x1 + 4]), (k, world, rank)
    with pytest.raises(ValueError):
        decomp.slab_bounds(90, 4, 0)
    assert decomp.neighbours(0, 4, True) == (3, 1) and decomp.neighbours(0, 4, False) == (None, 1)
    assert decomp.neighbours(3, 4, True) == (2, 0) and decomp.neighbours(3, 4, False) == (2, None)


def test_oracle_slab_invariance():
    """The property the multi-GPU design rests on: a slab computed with its 2-cell
    halos reproduces the global interior bit for bit -- checked here with the CPU oracle."""
    from oracle import oracle
    from veros_b200 import synthetic

    full = synthetic.make_workload("global_4deg", nx=24, ny=16)
    ref = oracle.isoneutral_step(copy_state(full))
    for x0, x1 in ((0, 12), (12, 24)):
        part = synthetic.make_workload("global_4deg", nx=x1 - x0, ny=16, x_offset=x0, nx_global=24)
        got = oracle.isoneutral_step(part)
        for k in ("temp", "salt", "dtemp_iso", "dsalt_iso", "P_diss_iso", "K_33", "Ai_ez", "Ai_by"):
            assert np.array_equal(got[k][2:-2], ref[k][x0 + 2:x1 + 2]), k


def test_vertmix_wrapper_validates_before_touching_the_library():
    """veros_b200.thermodynamics.vertmix_tempsalt: missing variables and CPU tensors are refused on the host,
    with the exception types of the reference's own op wrappers (tdma_.py:53-57); a state that carries only the
    vertmix variables is accepted by IsoState.from_numpy(strict=False)."""
    import helpers
    from veros_b200 import thermodynamics
    from veros_b200.state import IsoState

    st, _ = helpers.load_vmix_golden(helpers.vmix_golden_names()[0])
    s = IsoState.from_numpy(st, "cpu", strict=False)
    assert tuple(s.variables.kappaH.shape) == tuple(st["kappaH"].shape)
    assert tuple(s.variables.forc_temp_surface.shape) == tuple(st["kappaH"].shape[:2])
    assert not hasattr(s.variables, "K_iso")
    with pytest.raises(RuntimeError, match="no CPU path"):
        thermodynamics.vertmix_tempsalt(s)
    del s.variables.forc_salt_surface
    with pytest.raises(ValueError, match="forc_salt_surface"):
        thermodynamics.vertmix_tempsalt(s)
    with pytest.raises(KeyError):  # the isoneutral path needs its own variables
        IsoState.from_numpy(st, "cpu")
    s = IsoState.from_numpy(st, "cpu", strict=False)
    s.variables.kappaH = s.variables.kappaH[:, :, :-1]
    with pytest.raises(ValueError, match="kappaH"):
        s.validate(strict=False)


def test_balanced_slab_bounds_equalise_cost_and_respect_min_width():
    """Unequal-width x-slabs cut by cost (bench.py strong scaling): contiguous, covering, each at least min_width wide,
    per-slab cost within a few planes' worth of the mean; the 0.25 degree grid's continents make equal widths differ
    by +-25 %."""
    from veros_b200 import decomp, synthetic

    costs = synthetic.analytic_plane_costs("global_025deg")
    assert costs.shape == (1440,) and np.all(costs > 0)
    for world in (2, 4, 8):
        b = decomp.balanced_slab_bounds(costs, world)
        assert b[0][0] == 0 and b[-1][1] == 1440 and all(b[r][1] == b[r + 1][0] for r in range(world - 1))
        assert all(x1 - x0 >= 8 for x0, x1 in b)
        share = np.array([costs[x0:x1].sum() for x0, x1 in b]) / costs.sum() * world
        assert np.abs(share - 1).max() < 0.02
    even = np.array([costs[r * 180:(r + 1) * 180].sum() for r in range(8)]) / costs.sum() * 8
    assert even.max() > 1.1 and even.min() < 0.8
    # degenerate inputs: uniform costs give (almost) equal widths, narrow grids fall back to min_width, too narrow raises
    assert decomp.balanced_slab_bounds(np.ones(64), 4) == [(0, 16), (16, 32), (32, 48), (48, 64)]
    lop = decomp.balanced_slab_bounds(np.r_[np.ones(8) * 100, np.ones(24)], 4)
    assert all(x1 - x0 >= 8 for x0, x1 in lop) and lop[-1][1] == 32
    with pytest.raises(ValueError):
        decomp.balanced_slab_bounds(np.ones(20), 4)
    # the plane costs are those of the state the generator produces (same bathymetry code path)
    st = synthetic.make_workload("global_4deg")
    wet = st["maskT"][2:-2, 2:-2].reshape(90, -1).sum(axis=1)
    assert np.allclose(synthetic.analytic_plane_costs("global_4deg"), wet + 0.34 * (40 * 15 - wet))


def test_advance_time_rotates_device_scalars_host_copies_and_views():
    """the time-level indices the kernels read (device scalars) and the ones the plumbing uses (host copies,
    also in sub-slab views) must change together.  Runs on CPU tensors: IsoState only needs them as buffers."""
    from veros_b200 import synthetic
    from veros_b200.state import IsoState

    st = synthetic.make_workload("global_4deg", nx=12, ny=6)
    gs = IsoState.from_numpy(st, "cpu")
    sub = gs.subslab(0, 8)
    seen = []
    for _ in range(4):
        vs = gs.variables
        assert int(vs.tau.item()) == vs.tau_host and int(vs.taup1.item()) == vs.taup1_host
        assert sub.variables.taup1_host == vs.taup1_host and sub.variables.tau is vs.tau  # views share the scalars
        assert len({vs.tau_host, vs.taup1_host, 3 - vs.tau_host - vs.taup1_host}) == 3
        seen.append((vs.tau_host, vs.taup1_host))
        gs.advance_time()
    assert seen[0] == (1, 2) and seen[1] == (2, 0) and seen[2] == (0, 1) and seen[3] == seen[0]  # veros.py: taum1, tau, taup1 = tau, taup1, taum1


def test_jax_glue_module_imports_without_jax_and_targets_are_exported():
    """jax_glue.py imports jax lazily (there is none in this image): the module itself must load, every XLA target
    name must map to a symbol the library exports, and install() must refuse to run without the JAX backend."""
    from veros_b200 import _lib, build, jax_glue

    build.build()
    L = _lib.lib()
    for target, symbol in jax_glue.TARGETS.items():
        assert hasattr(L, symbol), (target, symbol)
    assert set(jax_glue.TARGETS.values()) >= set(_lib.OPS)
    assert jax_glue.build_tridiag_descriptor(3, 5) == bytes(_lib.TridiagDescriptor(num_systems=3, system_depth=5))
