"""Synthetic states for tests and benchmarks (NumPy, host side).

Every generator returns a "state dict" keyed by the reference's variable and setting names
(veros/variables.py, veros/settings.py) holding everything the isoneutral path reads:
    temp, salt, int_drhodT, int_drhodS (N,M,nz,3); K_iso, K_gm, dtemp_iso, dsalt_iso, P_diss_iso,
    P_diss_skew, K_11, K_22, K_33 (N,M,nz); Ai_* (N,M,nz,2,2); maskT/U/V/W (bool); kbot (int32);
    dxt, dxu (N); dyt, dyu, cost, cosu (M); dzt, dzw, zt (nz); tau, taup1; the settings.
N = nx + 4 and M = ny + 4 include two ghost cells per side (veros/variables.py:81,146-149).

* random_state: the stress state of the reference's unit tests (every field N(0,1), salt 35+N(0,1),
  random kbot with islands; veros/pyom_compat.py:445-526 describes the recipe) -- slopes are far
  beyond iso_slopec almost everywhere, masks and kbot are irregular.
* analytic_state: a smooth, stably stratified ocean on a spherical grid with bathymetry, continents
  and islands -- slopes of 1e-4..1e-2 so the taper works in its sensitive range.  Used for the
  global_4deg / global_1deg / 0.25 degree shaped benchmarks (the real setups need forcing files).
"""
import numpy as np

DEGTOM = 6370.0e3 * np.pi / 180.0  # veros/settings.py degtom (radius / 180 * pi)


def masks_from_kbot(kbot, nz, cyclic_x):
    """Mask rules of veros/core/numerics.py:200-221 (maskT from kbot, U/V/W as minima of neighbours)."""
    k = np.arange(nz)[None, None, :]
    maskT = (kbot > 0)[..., None] & (kbot[..., None] - 1 <= k)

    def wrap(m):
        if cyclic_x:
            m[-2:] = m[2:4]
            m[:2] = m[-4:-2]
        return m

    maskT = wrap(maskT)
    maskU = maskT.copy()
    maskU[:-1] = maskT[:-1] & maskT[1:]
    maskU = wrap(maskU)
    maskV = maskT.copy()
    maskV[:, :-1] = maskT[:, :-1] & maskT[:, 1:]
    maskV = wrap(maskV)
    maskW = maskT.copy()
    maskW[:, :, :-1] = maskT[:, :, :-1] & maskT[:, :, 1:]
    return maskT, maskU, maskV, maskW


def _finish(st, N, M, nz, settings):
    n3 = (N, M, nz)
    for k in ("dtemp_iso", "dsalt_iso", "P_diss_iso", "P_diss_skew"):
        st.setdefault(k, np.zeros(n3))
    for k in ("K_11", "K_22", "K_33"):
        st.setdefault(k, np.zeros(n3))
    for k in ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by"):
        st.setdefault(k, np.zeros(n3 + (2, 2)))
    st.update(settings)
    st["nx"], st["ny"], st["nz"] = N - 4, M - 4, nz
    return st


def random_state(nx, ny, nz, seed=17, eq_of_state_type=1, enable_cyclic_x=False, enable_conserve_energy=True,
                 dt_tracer=3600.0, K_iso_steep=1.0, iso_slopec=1e-3, iso_dslope=8e-4):
    """Settings default to test/pyom_consistency/isoneutral_test.py:9-19 on veros/settings.py defaults."""
    rng = np.random.default_rng(seed)
    N, M = nx + 4, ny + 4
    n3 = (N, M, nz)
    st = {}
    for name in ("temp", "salt", "int_drhodT", "int_drhodS"):
        st[name] = rng.standard_normal(n3 + (3,))
    st["salt"] += 35.0
    for name in ("K_iso", "K_gm", "dtemp_iso", "dsalt_iso", "P_diss_iso", "P_diss_skew", "K_11", "K_22", "K_33"):
        st[name] = rng.standard_normal(n3)
    for name in ("Ai_ez", "Ai_nz", "Ai_bx", "Ai_by"):
        st[name] = rng.standard_normal(n3 + (2, 2))

    kbot = np.zeros((N, M), dtype=np.int32)
    kbot[2:-2, 2:-2] = rng.integers(1, nz, size=(nx, ny))
    if nx > 2 and ny > 2:
        inner = kbot[3:-3, 3:-3]
        inner.flat[rng.integers(0, inner.size, size=10)] = 0
    if enable_cyclic_x:
        kbot[-2:] = kbot[2:4]
        kbot[:2] = kbot[-4:-2]
    st["kbot"] = kbot
    st["maskT"], st["maskU"], st["maskV"], st["maskW"] = masks_from_kbot(kbot, nz, enable_cyclic_x)

    def spacing(n, total):
        return total / n * (1.0 + 1e-2 * rng.standard_normal(n))

    dxt, dyt, dzt = spacing(N, 10_000e3), spacing(M, 10_000e3), spacing(nz, 6000.0 * nz / max(nx, 1))
    st["dxt"], st["dyt"], st["dzt"] = dxt, dyt, dzt
    st["dxu"] = 0.5 * (dxt + np.roll(dxt, -1))
    st["dyu"] = 0.5 * (dyt + np.roll(dyt, -1))
    st["dzw"] = 0.5 * (dzt + np.roll(dzt, -1))
    zw = np.cumsum(dzt) - dzt.sum()
    st["zt"] = zw - 0.5 * dzt
    st["cost"] = np.ones(M)
    st["cosu"] = np.ones(M)
    st["tau"], st["taup1"] = 1, 2
    settings = dict(eq_of_state_type=eq_of_state_type, enable_conserve_energy=enable_conserve_energy,
                    enable_cyclic_x=enable_cyclic_x, K_iso_steep=K_iso_steep, iso_slopec=iso_slopec,
                    iso_dslope=iso_dslope, dt_tracer=dt_tracer, grav=9.81, rho_0=1024.0)
    return _finish(st, N, M, nz, settings)


def vertical_grid(nz, depth=5000.0, dz_top=10.0):
    """Stretched levels, index 0 = bottom: dzt grows geometrically from dz_top at the surface."""
    if nz * dz_top >= depth:
        dz = np.full(nz, depth / nz)
    else:
        lo, hi = 1.0, 2.0
        for _ in range(200):  # growth factor r with dz_top * (r^nz - 1)/(r - 1) = depth
            r = 0.5 * (lo + hi)
            if dz_top * (r**nz - 1.0) / (r - 1.0) > depth:
                hi = r
            else:
                lo = r
        dz = dz_top * r ** np.arange(nz)
    dzt = dz[::-1].copy()
    zw = np.cumsum(dzt) - dzt.sum()  # top of each cell, zw[-1] = 0
    zt = zw - 0.5 * dzt
    dzw = np.empty(nz)
    dzw[:-1] = zt[1:] - zt[:-1]
    dzw[-1] = 0.5 * dzt[-1]
    return dzt, dzw, zt


def _analytic_kbot(ig, X2, Y2, dzt, nz, nxg, M, rng):
    """Bathymetry of the analytic ocean: ridges and basins, continents near fixed longitudes, a few random islands.
    `ig`: global interior index of every local column; X2 / Y2: longitude / latitude (radians) broadcastable to (N, M)."""
    depth = 5000.0 * (0.55 + 0.25 * np.cos(2 * X2) * np.cos(3 * Y2) + 0.2 * np.sin(5 * X2 + Y2))
    land = (np.cos(X2 - 0.6) > 0.93) & (np.abs(Y2) < 1.0)
    land |= (np.cos(X2 - 3.5) > 0.95) & (Y2 > -0.6)
    cell_top = np.cumsum(dzt) - dzt.sum()  # zw
    # kbot-1 = first level (from the bottom) whose top lies above the sea floor
    kb = 1 + (cell_top[None, None, :] <= -depth[..., None]).sum(axis=-1)
    kbot = np.clip(kb, 1, nz - 1).astype(np.int32)
    kbot[land] = 0
    jj, ii = rng.integers(4, max(5, M - 4), 12), rng.integers(0, nxg, 12)
    for a, b in zip(ii, jj):
        kbot[(ig == a), b] = 0
    kbot[:, :2] = 0
    kbot[:, -2:] = 0
    return kbot


def analytic_plane_costs(name, dry_cost=0.34, **overrides):
    """Relative cost of every interior x-plane of a named analytic workload, for load-balanced slab cuts: wet cells
    count 1, dry cells `dry_cost` (masked faces only store zeros and dry cells are skipped by the update, measured on
    B200: profiles/r02_scaling.md).  Only the 2-D bathymetry is generated (milliseconds, any grid size)."""
    cfg = dict(WORKLOADS[name])
    cfg.update(overrides)
    nxg, ny, nz = cfg["nx"], cfg["ny"], cfg["nz"]
    M = ny + 4
    lat_south, lat_north = cfg.get("lat_south", -78.0), cfg.get("lat_north", 78.0)
    rng = np.random.default_rng(cfg.get("seed", 0))
    ig = np.arange(nxg)
    lon = (ig + 0.5) * (360.0 / nxg)
    lat = lat_south + (np.arange(M) - 2 + 0.5) * ((lat_north - lat_south) / ny)
    dzt, _, _ = vertical_grid(nz)
    kbot = _analytic_kbot(ig, np.deg2rad(lon)[:, None], np.deg2rad(lat)[None, :], dzt, nz, nxg, M, rng)
    wet = np.where(kbot > 0, nz - (kbot - 1), 0)[:, 2:-2].sum(axis=1).astype(np.float64)  # wet cells per plane
    total = float(ny * nz)
    return wet + dry_cost * (total - wet)


def analytic_state(nx, ny, nz, eq_of_state_type=5, enable_conserve_energy=True, dt_tracer=86400.0 / 2,
                   K_iso_0=1000.0, K_iso_steep=500.0, iso_slopec=1e-3, iso_dslope=1e-3, lat_south=-78.0,
                   lat_north=78.0, seed=0, x_offset=0, nx_global=None):
    """Smooth stratified ocean on a lon/lat grid (cyclic in x).  `x_offset`/`nx_global` generate the
    x-slab [x_offset, x_offset + nx) of a wider global grid with identical values (multi-GPU)."""
    nxg = nx_global or nx
    N, M = nx + 4, ny + 4
    rng = np.random.default_rng(seed)
    dlon = 360.0 / nxg
    dlat = (lat_north - lat_south) / ny
    ig = (np.arange(N) - 2 + x_offset) % nxg  # global interior index of every local column
    lon = (ig + 0.5) * dlon
    lat = lat_south + (np.arange(M) - 2 + 0.5) * dlat
    latu = lat + 0.5 * dlat
    dzt, dzw, zt = vertical_grid(nz)
    st = dict(dzt=dzt, dzw=dzw, zt=zt)
    st["dxt"] = np.full(N, dlon * DEGTOM)
    st["dxu"] = np.full(N, dlon * DEGTOM)
    st["dyt"] = np.full(M, dlat * DEGTOM)
    st["dyu"] = np.full(M, dlat * DEGTOM)
    st["cost"] = np.cos(np.deg2rad(lat))
    st["cosu"] = np.cos(np.deg2rad(latu))

    X = np.deg2rad(lon)[:, None, None]
    Y = np.deg2rad(lat)[None, :, None]
    Z = zt[None, None, :]
    # thermocline that shoals towards the poles and undulates zonally -> sloping isopycnals
    hT = 300.0 + 700.0 * np.cos(Y) ** 2 * (1.0 + 0.15 * np.sin(3 * X) * np.cos(2 * Y))
    temp = -1.0 + 26.0 * np.cos(Y) ** 2 * np.exp(Z / hT) + 2.0 * np.exp(Z / 2500.0) + 0.3 * np.sin(5 * X + 3 * Y) * np.exp(Z / 600.0)
    salt = 34.7 + 0.8 * np.cos(2 * Y) * np.exp(Z / 400.0) + 0.15 * np.cos(4 * X - 2 * Y) * np.exp(Z / 900.0)
    n3 = (N, M, nz)
    st["temp"] = np.empty(n3 + (3,))
    st["salt"] = np.empty(n3 + (3,))
    for lev, eps in enumerate((-1.0, 0.0, 1.0)):  # three slightly different time levels
        st["temp"][..., lev] = temp * (1.0 + 1e-4 * eps) + 1e-3 * eps * np.sin(2 * X) * np.exp(Z / 300.0)
        st["salt"][..., lev] = salt + 1e-4 * eps * np.cos(3 * X + Y)
    st["K_iso"] = K_iso_0 * (1.0 + 0.3 * np.cos(2 * X) * np.cos(Y) * np.exp(Z / 1500.0)) * np.ones(n3)
    st["K_gm"] = 0.8 * st["K_iso"]
    rho0, grav, betaT, betaS = 1024.0, 9.81, 1.67e-4, 0.78e-3
    st["int_drhodT"] = np.repeat((-rho0 * betaT * Z * (1.0 + 0.02 * temp))[..., None], 3, axis=-1)
    st["int_drhodS"] = np.repeat((rho0 * betaS * Z * (1.0 + 0.001 * (salt - 35.0)))[..., None], 3, axis=-1)

    kbot = _analytic_kbot(ig, X[..., 0], Y[..., 0], dzt, nz, nxg, M, rng)
    st["kbot"] = kbot
    st["maskT"], st["maskU"], st["maskV"], st["maskW"] = masks_from_kbot(kbot, nz, cyclic_x=False)
    st["tau"], st["taup1"] = 1, 2
    settings = dict(eq_of_state_type=eq_of_state_type, enable_conserve_energy=enable_conserve_energy,
                    enable_cyclic_x=True, K_iso_steep=K_iso_steep, iso_slopec=iso_slopec, iso_dslope=iso_dslope,
                    dt_tracer=dt_tracer, grav=grav, rho_0=rho0)
    return _finish(st, N, M, nz, settings)


# Shapes of the BASELINE.json configs (interior cells)
WORKLOADS = {
    "acc": dict(nx=30, ny=42, nz=15, eq_of_state_type=3, iso_slopec=0.01, iso_dslope=0.005),
    "bench_1M": dict(nx=142, ny=142, nz=50),                       # run_benchmarks.py:188-192, --sizes 1e6
    "global_4deg": dict(nx=90, ny=40, nz=15, eq_of_state_type=5),
    "global_1deg": dict(nx=360, ny=160, nz=115, eq_of_state_type=5),
    "global_025deg": dict(nx=1440, ny=720, nz=80, eq_of_state_type=5),
}


def make_workload(name, nx=None, x_offset=0, nx_global=None, **overrides):
    """State dict for a named workload; `bench_1M` is the random stress state, the rest analytic."""
    cfg = dict(WORKLOADS[name])
    cfg.update(overrides)
    if nx is not None:
        cfg["nx"] = nx
    if name == "bench_1M":
        return random_state(cfg.pop("nx"), cfg.pop("ny"), cfg.pop("nz"), **cfg)
    return analytic_state(cfg.pop("nx"), cfg.pop("ny"), cfg.pop("nz"), x_offset=x_offset,
                          nx_global=nx_global, **cfg)
