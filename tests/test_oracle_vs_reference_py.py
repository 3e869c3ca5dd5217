"""The oracle's ray tracer against the REFERENCE'S OWN PYTHON, bit for bit (CPU; runs where /root/reference exists).

``oracle/refpy.py`` imports ``RAiDER.delay``, ``RAiDER.losreader``, ``RAiDER.delayFcns`` and ``RAiDER.utilFcns`` unmodified
from ``/root/reference/tools/RAiDER`` (stand-ins only for pyproj / xarray / rasterio / shapely, which are not installable
offline; the pyproj stand-in computes with ``oracle.geodesy``).  Every test below runs the reference function and the
restatement in ``oracle/raytrace.py`` on the same inputs and asserts ``np.array_equal`` -- so the loop structure, the
global ``nParts``, the ``.all()`` clamps, the Newton schedule, the trapezoid weights and the NumPy rounding order of
``delay.py:196-326`` and ``losreader.py:706-733,772-835`` are pinned to the reference itself, and so are the committed
golden vectors (``tests/golden/raytrace.npz`` is written from the reference functions' outputs by ``make_golden.py``).

On the GPU box ``/root/reference`` is absent: these tests skip there and the goldens stand in.
"""
import numpy as np
import pytest

from oracle import geodesy, raytrace as rt, refpy
from raider_b200 import synthetic as syn

pytestmark = pytest.mark.skipif(not refpy.available(), reason='/root/reference is only present in the build container')


@pytest.fixture(scope='module')
def ref():
    return refpy.load()


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def _rays(n=96, seed=3, lat=(-75, 75), hmax=3000.0):
    rng = np.random.default_rng(seed)
    la, lo, h = rng.uniform(*lat, n), rng.uniform(-180, 180, n), rng.uniform(-100, hmax, n)
    enu = geodesy.inc_hd_to_enu(rng.uniform(0, 65, n), rng.uniform(0, 360, n))
    look = geodesy.enu2ecef(enu[:, 0], enu[:, 1], enu[:, 2], la, lo, h)
    g = np.stack(geodesy.lla2ecef(la, lo, h), axis=-1)
    return la, lo, h, g.reshape(n // 8, 8, 3), look.reshape(n // 8, 8, 3), enu[:, 2].reshape(n // 8, 8)


def test_reference_modules_are_the_reference(ref):
    """What is imported is the file under /root/reference, not a copy, and pyproj is a stand-in (so say the pins)."""
    for m in (ref.delay, ref.losreader, ref.delayFcns, ref.utilFcns):
        assert m.__file__.startswith(refpy.REFERENCE_ROOT)
    import pyproj
    assert getattr(pyproj, '__raider_b200_standin__', False)


def test_geodesy_helpers_bitwise(ref):
    la, lo, h, g, look, cosf = _rays()
    assert same(np.stack(ref.utilFcns.lla2ecef(la, lo, h), -1), np.stack(geodesy.lla2ecef(la, lo, h), -1))
    x, y, z = g[..., 0], g[..., 1], g[..., 2]
    assert same(np.stack(ref.utilFcns.ecef2lla(x, y, z), -1), np.stack(geodesy.ecef2lla(x, y, z), -1))
    inc, hd = np.linspace(0, 70, 29), np.linspace(-180, 360, 29)
    assert same(ref.losreader.inc_hd_to_enu(inc, hd), geodesy.inc_hd_to_enu(inc, hd))
    enu = geodesy.inc_hd_to_enu(inc, hd)
    assert same(ref.utilFcns.enu2ecef(enu[:, 0], enu[:, 1], enu[:, 2], la[:29], lo[:29], h[:29]),
                geodesy.enu2ecef(enu[:, 0], enu[:, 1], enu[:, 2], la[:29], lo[:29], h[:29]))
    assert same(ref.losreader.getZenithLookVecs(la, lo, h), geodesy.getZenithLookVecs(la, lo, h))
    assert same(ref.utilFcns.ecef2enu(look, la.reshape(-1, 8), lo.reshape(-1, 8), 0.0),
                geodesy.ecef2enu(look, la.reshape(-1, 8), lo.reshape(-1, 8)))
    with pytest.raises(ValueError):
        ref.losreader.inc_hd_to_enu(-1.0, 0.0)
    with pytest.raises(ValueError):
        geodesy.inc_hd_to_enu(-1.0, 0.0)


@pytest.mark.parametrize('toa', [0.0, 1234.5, 26000.0, 48000.0, 80000.0])
def test_get_top_of_atmosphere_bitwise(ref, toa):
    _, _, _, g, look, cosf = _rays(seed=int(toa) % 97)
    assert same(ref.losreader.getTopOfAtmosphere(g, look, toa), rt.getTopOfAtmosphere(g, look, toa))
    assert same(ref.losreader.getTopOfAtmosphere(g, look, toa, factor=cosf), rt.getTopOfAtmosphere(g, look, toa, factor=cosf))


@pytest.mark.parametrize('table,ht,zref', [
    ('nz37', 0.0, None), ('nz37', 1500.0, 26000.0), ('nz37', -499.5, 100.0), ('nz37', 46000.0, None),
    ('ml145', 0.0, None), ('ml145', 733.0, 26000.0), ('hrrr57', 200.0, None), ('hrrr57', -500.0, 15000.0),
])
def test_build_ray_bitwise(ref, table, ht, zref):
    zs = syn.z_levels(37) if table == 'nz37' else syn.z_levels_table(table)
    zref = zs[-1] - 1 if zref is None else zref
    _, _, _, g, look, _ = _rays(seed=11, hmax=0.0)
    a = ref.losreader.build_ray(zs, ht, g, look, zref)
    b = rt.build_ray(zs, ht, g, look, zref)
    assert a[0] is not None
    for u, v in zip(a, b):
        assert same(u, v)
    # the scalar layer decisions (losreader.py:785-809) the device plan is built from
    assert len(rt.layer_plan(zs, ht, zref)) == a[0].shape[0]


def test_build_ray_no_contributing_layer(ref):
    zs = syn.z_levels(37)
    _, _, _, g, look, _ = _rays(seed=2)
    assert ref.losreader.build_ray(zs, zs[-1], g, look, zs[-1] - 1) == (None, None, None)
    assert rt.build_ray(zs, zs[-1], g, look, zs[-1] - 1) == (None, None, None)
    # layers thinner than 1 m are skipped (losreader.py:808)
    assert ref.losreader.build_ray(np.array([0.0, 0.5, 0.9]), 0.0, g, look, 10.0) == (None, None, None)
    assert rt.build_ray(np.array([0.0, 0.5, 0.9]), 0.0, g, look, 10.0) == (None, None, None)


def _ref_crs(ref, crs):
    if crs is None or crs.kind == 0:
        return ref.CRS.from_epsg(4326)
    l = crs.lcc
    return ref.CRS(dict(proj='lcc', lat_1=l.lat_1, lat_2=l.lat_2, lat_0=l.lat_0, lon_0=l.lon_0, a=l.R, b=l.R, x_0=l.x_0, y_0=l.y_0))


def _ref_los(ref, los):
    if isinstance(los, rt.ZenithLOS):
        return refpy.zenith_los(ref)
    if isinstance(los, rt.FixedIncidenceLOS):
        return refpy.fixed_incidence_los(ref, los.incidence_deg, los.heading_deg)
    return refpy.array_los(los.vecs)


def _both(ref, cfg, los, model_crs=None, pts_crs=None, output=None):
    zpts = np.asarray(cfg['zpts'], dtype=np.float64)
    ifs_ref = ref.delayFcns.getInterpolators(refpy.dataset(cfg['cube']), 'pointwise')
    ifs = rt.get_interpolators(cfg['cube'], 'pointwise')
    for a, b in zip(ifs_ref, ifs):       # a3: same grid tuple (ys, xs, zs), same fp32 values in the transposed view
        assert all(same(u, v) for u, v in zip(a.grid, b.grid)) and same(a.values, b.values)
    kw = dict(MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
    if output is not None:
        o_ref, o_port = [x.copy() for x in output], [x.copy() for x in output]
        assert ref.delay._build_cube_ray(cfg['xpts'], cfg['ypts'], zpts, _ref_los(ref, los), _ref_crs(ref, model_crs),
                                         _ref_crs(ref, pts_crs), list(ifs_ref), outputArrs=o_ref, **kw) is None
        assert rt.build_cube_ray(cfg['xpts'], cfg['ypts'], zpts, los, model_crs or rt.GeographicCRS(), pts_crs or rt.GeographicCRS(),
                                 list(ifs), outputArrs=o_port, **kw) is None
        return o_ref, o_port
    o_ref = ref.delay._build_cube_ray(cfg['xpts'], cfg['ypts'], zpts, _ref_los(ref, los), _ref_crs(ref, model_crs),
                                      _ref_crs(ref, pts_crs), list(ifs_ref), **kw)
    o_port = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], zpts, los, model_crs or rt.GeographicCRS(), pts_crs or rt.GeographicCRS(),
                               list(ifs), **kw)
    return o_ref, o_port


def _golden_cases():
    """The six geometries of tests/golden/make_golden.py::golden_raytrace (a-d, f; e is the zenith cube below)."""
    cfg = syn.config_c2(n=24)
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 24, 24, 0.08)
    yield 'a', cfg, rt.FixedIncidenceLOS(cfg['incidence'], cfg['heading']), None
    cfg = syn.config_c2(n=12, table='ml145')
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 12, 12, 0.15)
    cfg['zpts'] = np.array([0.0, 1500.0])
    yield 'b', cfg, rt.FixedIncidenceLOS(45.0, 12.0), None
    cfg = syn.config_c2(n=10)
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 10, 10, 0.2)
    yield 'c', cfg, rt.ZenithLOS(), None
    cfg = syn.config_c2(n=16)
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 16, 16, 0.1)
    xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
    inc = 20.0 + 26.0 * (xx - xx.min()) / (xx.max() - xx.min())
    enu = geodesy.inc_hd_to_enu(inc, np.full(inc.shape, -168.0))
    yield 'd', cfg, rt.ArrayLOS(geodesy.enu2ecef(enu[..., 0], enu[..., 1], enu[..., 2], yy, xx, 0 * yy)), None
    lcc = rt.LambertCRS()
    cx, cy = lcc.lcc.forward(-98.0, 36.0)
    xs = cx + 3000.0 * (np.arange(60) - 30)
    ys = cy + 3000.0 * (np.arange(50) - 25)
    X, Y = np.meshgrid(xs, ys)
    lon_n, lat_n = lcc.lcc.inverse(X, Y)
    cube = syn.make_cube(ys, xs, syn.z_levels_table('hrrr57'), lat_of=lat_n, lon_of=lon_n, seed=5)
    xpts, ypts = syn.raster(36.0, -98.0, 12, 12, 0.03)
    yield 'f', dict(cube=cube, xpts=xpts, ypts=ypts, zpts=np.array([200.0]), zref=float(cube['z'][-1] - 1),
                    max_segment_length=1000.0), rt.FixedIncidenceLOS(35.0, -12.0), lcc


@pytest.mark.parametrize('case', list('abcdf'))
def test_build_cube_ray_bitwise_and_golden(ref, golden, case):
    """reference _build_cube_ray == oracle port == committed golden, bit for bit, on the golden geometries."""
    g = golden('raytrace')
    name, cfg, los, crs = next(c for c in _golden_cases() if c[0] == case)
    o_ref, o_port = _both(ref, cfg, los, model_crs=crs)
    assert same(o_ref[0], o_port[0]) and same(o_ref[1], o_port[1])
    assert same(o_ref[0], g[f'{case}_wet']) and same(o_ref[1], g[f'{case}_hydro'])
    assert np.isfinite(o_ref[0]).all() and (o_ref[1] > 0).all()


def test_build_cube_ray_accumulates_in_place_bitwise(ref):
    name, cfg, los, crs = next(c for c in _golden_cases() if c[0] == 'c')
    rng = np.random.default_rng(8)
    seed = [rng.normal(size=(1, 10, 10)), rng.normal(size=(1, 10, 10))]
    o_ref, o_port = _both(ref, cfg, los, output=seed)
    assert same(o_ref[0], o_port[0]) and same(o_ref[1], o_port[1])
    assert not same(o_ref[0], seed[0])


def test_build_cube_ray_edge_rules_bitwise(ref):
    """Rays leaving the cube sideways (NaN pixels), the top output slice that is skipped (zeros), the first-sample
    clamp of delay.py:306-307 on a raster that sits exactly on the lowest node, and a projected (Lambert) query grid."""
    # (1) raster hanging over the cube edge: NaN where the ray leaves, identical NaN pattern
    cfg = syn.config_c2(n=8)
    cfg['xpts'], cfg['ypts'] = syn.raster(float(cfg['cube']['y'][-1]) - 0.05, -118.0, 8, 8, 0.02)
    o_ref, o_port = _both(ref, cfg, rt.FixedIncidenceLOS(40.0, -168.0))
    assert same(o_ref[0], o_port[0]) and same(o_ref[1], o_port[1])
    assert np.isnan(o_ref[0]).any() and np.isfinite(o_ref[0]).any()
    # (2) zpts ending at the model top: last slice stays zero (delay.py:276-277)
    cfg = syn.config_c2(n=6)
    cfg['zpts'] = np.array([0.0, float(cfg['cube']['z'][-1])])
    o_ref, o_port = _both(ref, cfg, rt.FixedIncidenceLOS(30.0, -168.0))
    assert same(o_ref[0], o_port[0]) and (o_ref[0][1] == 0).all() and (o_ref[0][0] > 0).all()
    # (3) output height == lowest node: Newton leaves the first sample a hair below min(z) for every pixel -> clamp
    cfg = syn.config_c2(n=6)
    cfg['zpts'] = np.array([float(cfg['cube']['z'][0])])
    o_ref, o_port = _both(ref, cfg, rt.FixedIncidenceLOS(30.0, -168.0))
    assert same(o_ref[0], o_port[0]) and same(o_ref[1], o_port[1])
    # (3b) the upper clamp (delay.py:310-311): at the default zref (1 m below the model top) the third Newton iterate overshoots the
    # top of an 80 km table by more than that metre from ~58 deg incidence on -- on every pixel, so the last sample is taken at max(z)
    from pathlib import Path
    zs = np.load(Path(__file__).resolve().parent / 'golden' / 'era5_slant_ref.npz')['z']
    xp, yp = syn.raster(33.5, -117.8, 5, 5, 0.02)
    xs, ys = syn.cube_axes_around(xp, yp, pad_deg=3.0)
    cfg = {'cube': syn.make_cube(ys, xs, zs, totals=False), 'xpts': xp, 'ypts': yp, 'zpts': np.array([0.0]), 'zref': float(zs[-1] - 1),
           'max_segment_length': 1000.0}
    o_ref, o_port = _both(ref, cfg, rt.FixedIncidenceLOS(66.0, -168.0))
    assert same(o_ref[0], o_port[0]) and same(o_ref[1], o_port[1]) and np.isfinite(o_ref[0]).all()
    g = np.stack(geodesy.lla2ecef(yp[0], xp[0], 0.0))
    u = np.asarray(rt.FixedIncidenceLOS(66.0, -168.0).getLookVectors(0.0, [xp[:1, None], yp[:1, None], np.zeros((1, 1))], None, None))[0, 0]
    top = ref.losreader.build_ray(zs, 0.0, g[None, None, :], u[None, None, :], cfg['zref'])[2][-1][0, 0]
    assert geodesy.ecef2height(top[0], top[1], top[2]) > zs[-1] + 1.0      # (the overshoot itself, on the reference's own build_ray)
    # (4) query raster given in the Lambert system of the cube (pts_crs != 4326: delay.py:262-265)
    name, cfgf, los, lcc = next(c for c in _golden_cases() if c[0] == 'f')
    cx, cy = lcc.lcc.forward(-98.0, 36.0)
    cfgp = dict(cfgf, xpts=cx + 2500.0 * np.arange(-4, 4), ypts=cy + 2500.0 * np.arange(-3, 3))
    o_ref, o_port = _both(ref, cfgp, los, model_crs=lcc, pts_crs=lcc)
    assert same(o_ref[0], o_port[0]) and same(o_ref[1], o_port[1]) and np.isfinite(o_ref[0]).all()


def test_build_cube_ray_error_rules(ref):
    """delay.py:276-280: a non-last height without layers -> TypeError (np.isnan(None)); all-NaN lengths -> ValueError."""
    cfg = syn.config_c2(n=4)
    top = float(cfg['cube']['z'][-1])
    ifs_ref = list(ref.delayFcns.getInterpolators(refpy.dataset(cfg['cube'])))
    ifs = list(rt.get_interpolators(cfg['cube']))
    crs = ref.CRS.from_epsg(4326)
    z_bad = np.array([top, 0.0])
    with pytest.raises(TypeError):
        ref.delay._build_cube_ray(cfg['xpts'], cfg['ypts'], z_bad, refpy.zenith_los(ref), crs, crs, ifs_ref, MAX_TROPO_HEIGHT=cfg['zref'])
    with pytest.raises(TypeError):
        rt.build_cube_ray(cfg['xpts'], cfg['ypts'], z_bad, rt.ZenithLOS(), rt.GeographicCRS(), rt.GeographicCRS(), ifs, MAX_TROPO_HEIGHT=cfg['zref'])
    nan_los = np.full((4, 4, 3), np.nan)
    with pytest.raises(ValueError, match='geo2rdr did not converge'):
        ref.delay._build_cube_ray(cfg['xpts'], cfg['ypts'], np.array([0.0]), refpy.array_los(nan_los), crs, crs, ifs_ref, MAX_TROPO_HEIGHT=cfg['zref'])
    with pytest.raises(ValueError, match='geo2rdr did not converge'):
        rt.build_cube_ray(cfg['xpts'], cfg['ypts'], np.array([0.0]), rt.ArrayLOS(nan_los), rt.GeographicCRS(), rt.GeographicCRS(), ifs, MAX_TROPO_HEIGHT=cfg['zref'])


def test_build_cube_zenith_bitwise_and_golden(ref, golden):
    g = golden('raytrace')
    c1 = syn.config_c1()
    crs = ref.CRS.from_epsg(4326)
    zt = ref.delay._build_cube(c1['xpts'][::5], c1['ypts'][::5], np.asarray(c1['zpts']), crs, crs,
                               list(ref.delayFcns.getInterpolators(refpy.dataset(c1['cube']), 'total')))
    gg = rt.GeographicCRS()
    zp = rt.build_cube(c1['xpts'][::5], c1['ypts'][::5], c1['zpts'], gg, gg, list(rt.get_interpolators(c1['cube'], 'total')))
    assert same(zt[0], zp[0]) and same(zt[1], zp[1])
    assert same(zt[0], g['e_wet']) and same(zt[1], g['e_hydro'])


def test_build_cube_projected_model_bitwise(ref):
    """_build_cube with model_crs != pts_crs goes through transformPoints (delay.py:404-436)."""
    name, cfgf, los, lcc = next(c for c in _golden_cases() if c[0] == 'f')
    cube = dict(cfgf['cube'])
    zs = np.array([0.0, 500.0])
    a = ref.delay._build_cube(cfgf['xpts'], cfgf['ypts'], zs, _ref_crs(ref, lcc), ref.CRS.from_epsg(4326),
                              list(ref.delayFcns.getInterpolators(refpy.dataset(cube), 'total')))
    b = rt.build_cube(cfgf['xpts'], cfgf['ypts'], zs, lcc, rt.GeographicCRS(), list(rt.get_interpolators(cube, 'total')))
    assert same(a[0], b[0]) and same(a[1], b[1]) and np.isfinite(a[0]).all()


def test_geodesy_golden_is_reference_output(ref, golden):
    g = golden('geodesy')
    assert same(ref.losreader.getTopOfAtmosphere(g['g0'], g['look'], 30000.0), g['toa10'])
    assert same(ref.losreader.getTopOfAtmosphere(g['g0'], g['look'], 30000.0, factor=g['cosf']), g['toa3'])
    lens, lows, highs = ref.losreader.build_ray(g['zs'], 0.0, g['g0'], g['look'], g['zs'][-1] - 1)
    assert same(lens, g['lens']) and same(lows, g['lows']) and same(highs, g['highs'])


def test_constant_refractivity_identity_on_the_reference(ref):
    """test/test_synthetic.py:217-274 restated on the reference's own _build_cube_ray: delay = N 1e-6 sum_k L_k."""
    cfg = syn.config_c2(n=6)
    cube = dict(cfg['cube'])
    cube['wet'] = np.full_like(cube['wet'], 40.0)
    cube['hydro'] = np.full_like(cube['hydro'], 250.0)
    cfg['cube'] = cube
    o_ref, o_port = _both(ref, cfg, rt.FixedIncidenceLOS(30.0, -168.0))
    xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
    xyz = np.stack(ref.utilFcns.lla2ecef(yy, xx, 0 * yy), -1)
    look = refpy.fixed_incidence_los(ref, 30.0, -168.0).getLookVectors(0.0, [xx, yy, 0 * yy], xyz, yy)
    lens, _, _ = ref.losreader.build_ray(cube['z'], 0.0, xyz, look, cfg['zref'])
    assert np.allclose(o_ref[0][0], 40e-6 * lens.sum(0), rtol=1e-12, atol=0)
    assert np.allclose(o_ref[1][0], 250e-6 * lens.sum(0), rtol=1e-12, atol=0)


def test_los_contract_matches_reference_classes(ref):
    """raider_b200.losreader's LOS duck type against the reference classes (losreader.py:32-91): predicates, setPoints forms."""
    from raider_b200 import losreader as mine
    rng = np.random.default_rng(0)
    lats, lons, hts = rng.uniform(-60, 60, (4, 5)), rng.uniform(-180, 180, (4, 5)), rng.uniform(0, 100, (4, 5))
    for cls_ref, cls_mine in [(ref.losreader.Zenith, mine.Zenith)]:
        a, b = cls_ref(), cls_mine()
        assert (a.is_Zenith(), a.is_Projected(), a.ray_trace()) == (b.is_Zenith(), b.is_Projected(), b.ray_trace())
        with pytest.raises(RuntimeError):
            a.setPoints(None)
        with pytest.raises(RuntimeError):
            b.setPoints(None)
        with pytest.raises(ValueError):
            a.setLookVectors()
        with pytest.raises(ValueError):
            b.setLookVectors()
        for args in [(lats, lons, hts), (lats, lons), (np.stack([lats, lons, hts], -1),)]:
            a.setPoints(*args)
            b.setPoints(*args)
            assert same(a._lats, b._lats) and same(a._lons, b._lons) and same(a._heights, b._heights)
        a._look_vecs = b._look_vecs = None
        a.setLookVectors()
        b.setLookVectors()
        assert same(a._look_vecs, b._look_vecs)
        d = rng.normal(size=(4, 5))
        assert same(a(d), b(d))
    # mode flags of the ray-tracing / projected classes
    r = mine.Raytracing(incidence=30.0, heading=-168.0)
    assert (r.is_Zenith(), r.is_Projected(), r.ray_trace()) == (False, False, True)
    c = mine.Conventional(incidence=30.0)
    assert (c.is_Zenith(), c.is_Projected(), c.ray_trace()) == (False, True, False)
    # Conventional.__call__ = delays / cos(inc) via inc_hd_to_enu's up component (losreader.py:110-123)
    c.setPoints(lats, lons, hts)
    d = rng.uniform(1, 3, size=(4, 5))
    assert same(c(d), d / ref.losreader.inc_hd_to_enu(np.float64(30.0), np.float64(0.0))[..., -1])


# ------------------------------------------------------------------- the reference's end-to-end goldens of the slant path
def test_reference_goldens_of_test_slant_are_reproduced(ref):
    """test/test_slant.py:49 (2.333865144 m, projected LOS on a cube = zenith totals) and :99 (2.97711681 m, Raytracing through a
    Sentinel-1 precise-orbit file) on the reference's own ERA-5 cube, to the 7 decimals the reference asserts: reference AOI
    grid + reference Raytracing / get_orbit / _build_cube_ray, with the NetCDF-4 file read by raider_b200.hdf5_lite and isce3
    replaced by oracle/orbit.py's restatement -- THE pin of that restatement.  The port (oracle.raytrace + oracle.orbit.OrbitLOS)
    is bitwise equal to the reference run, and the committed fixture (tests/golden/era5_slant_ref.npz) holds exactly these
    arrays.  One height only here (8100 Newton solves in pure Python); make_golden_era5_slant.py ran all four."""
    import sys
    from pathlib import Path
    from oracle import orbit as ob
    sys.path.insert(0, str(Path(__file__).resolve().parent / 'golden'))
    import make_golden_era5_slant as mk
    fx = np.load(Path(__file__).resolve().parent / 'golden' / 'era5_slant_ref.npz')
    _, cube, aoi = mk.reference_setup()
    assert np.array_equal(aoi.xpts, fx['xpts']) and np.array_equal(aoi.ypts, fx['ypts'])
    for k in ('x', 'y', 'z', 'wet', 'hydro', 'wet_total', 'hydro_total'):
        assert same(cube[k], fx[k])
    iy, ix = mk.gold_index(aoi)
    assert (iy, ix) == tuple(fx['gold_index'])
    zw, zh = mk.reference_ztd(ref, cube, aoi, fx['zpts'])
    np.testing.assert_almost_equal(mk.GOLD_STD, (zw + zh)[0, iy, ix])
    assert same(zw, fx['ref_ztd_wet']) and same(zh, fx['ref_ztd_hydro'])
    z0 = fx['zpts'][:1]
    (rw, rh), los, toa = mk.reference_ray(ref, cube, aoi, z0)
    np.testing.assert_almost_equal(mk.GOLD_RAY, (rw + rh)[0, iy, ix])
    assert toa == float(fx['zref']) and same(rw[0], fx['ref_ray_wet'][0]) and same(rh[0], fx['ref_ray_hydro'][0])
    o = los._orbit
    crs = rt.GeographicCRS()
    pw, ph = rt.build_cube_ray(aoi.xpts, aoi.ypts, z0, ob.OrbitLOS(ob.Orbit(o.time, o.position, o.velocity)), crs, crs,
                               list(rt.get_interpolators(cube)), MAX_TROPO_HEIGHT=toa)
    assert same(pw, rw) and same(ph, rh)
    # the state vectors of the fixture's text file are the ones get_sv kept
    from raider_b200.losreader import read_txt_file
    sv = read_txt_file(str(Path(__file__).resolve().parent / 'golden' / 'orbit_S1B_20200130_sv.txt'))
    assert np.array_equal(np.stack(sv[1:4], -1), o.position) and np.array_equal(np.stack(sv[4:7], -1), o.velocity)


def test_reference_golden_of_test_gnss_intersect_is_reproduced(ref):
    """test/test_intersect.py:104 (TORP 2.34514 m, 4 decimals): the station (point) mode of tropo_delay on the reference's ERA-5 file --
    llreader.StationFile (pandas) -> ZTD cube on the AOI grid at the model's 145 z levels -> getInterpolators(ds, 'ztd') at the stations
    (delay.py:78-121) -- run with the reference's own Python; the oracle (build_cube + the installed scipy) gives the same numbers, and
    the committed fixture (tests/golden/era5_gnss_ref.npz) holds them for the GPU box."""
    import sys
    from pathlib import Path
    from scipy.interpolate import RegularGridInterpolator as RGI
    sys.path.insert(0, str(Path(__file__).resolve().parent / 'golden'))
    import make_golden_era5_slant as mk
    from raider_b200.cube_io import load_cube
    fx = np.load(Path(__file__).resolve().parent / 'golden' / 'era5_gnss_ref.npz')
    aoi, (lats, lons, hgts), (zw, zh), (wd, hd) = mk.reference_station_ztd()
    np.testing.assert_almost_equal((wd + hd)[int(fx['gold_index'])], float(fx['gold_total']), decimal=4)
    assert same(aoi.xpts, fx['xpts']) and same(aoi.ypts, fx['ypts']) and same(wd, fx['ref_wet']) and same(hd, fx['ref_hydro'])
    assert same(lats, fx['lats']) and same(lons, fx['lons']) and same(hgts, fx['hgts'])
    cube = load_cube(mk.WM)
    crs = rt.GeographicCRS()
    pw, ph = rt.build_cube(aoi.xpts, aoi.ypts, cube['z'], crs, crs, list(rt.get_interpolators(cube, 'total')))
    assert same(pw, zw) and same(ph, zh)
    pts = np.stack([lats, lons, hgts], axis=-1)
    grid = (aoi.ypts, aoi.xpts, cube['z'])
    assert same(RGI(grid, pw.transpose(1, 2, 0), fill_value=np.nan, bounds_error=False)(pts), wd)
    assert same(RGI(grid, ph.transpose(1, 2, 0), fill_value=np.nan, bounds_error=False)(pts), hd)
