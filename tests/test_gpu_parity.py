"""GPU parity tests proper: CUDA path (through the C ABI / the RAiDER-shaped shims) vs the oracle and the golden vectors.

Tolerances (BASELINE.json north_star): bit-exact for the integer ray-step counts (nParts) and for the kernels whose
arithmetic is pinned op-for-op (makePoints, RAiDER.interpolate, interpolate_along_axis, the scipy-order trilinear
sampler); |delta| <= 1e-6 m for fp64 integrated delays (observed ~1e-12); 1e-3 m for the fp32 output tier.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_F64_M = 1e-6
TOL_F32_M = 1e-3


@pytest.fixture(scope='module')
def gpu(lib):
    if lib.rdr_device_count() < 1:
        pytest.fail('no CUDA device visible to libraider_b200.so on a box that runs -m gpu tests')
    return lib


def _c2_small(n, posting, **kw):
    from raider_b200 import synthetic as syn
    cfg = syn.config_c2(n=n, **kw)
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, n, n, posting)
    return cfg


# ---------------------------------------------------------------------------------------- K1
def test_makepoints_bit_exact(gpu, golden):
    from raider_b200.makePoints import makePoints0D, makePoints1D, makePoints2D, makePoints3D
    g = golden('makepoints')
    out = makePoints3D(100.0, g['sp3'], g['slv3'], 5)
    assert out.ndim == 5 and np.array_equal(out, g['out3'])           # test_result_makePoints3D.txt, bit for bit
    assert np.array_equal(makePoints2D(5000.0, g['sp2'], g['slv2'], 15.0), g['out2'])
    for L, s, n in g['counts']:
        assert makePoints0D(L, np.zeros(3), np.ones(3), s).shape == (3, int(n))
    # fixtures of test/test_util.py:49-89
    ray = makePoints0D(1000.0, np.array([0.0, 0.0, 0.0]), np.array([0.0, 0.0, 1.0]), 5.0)
    assert np.allclose(ray, np.stack([np.zeros(200), np.zeros(200), np.arange(0, 1000, 5)], axis=-1).T)
    sp = np.zeros((2, 3)); slv = np.array([[0.0, 0, 1], [0, 1.0, 0]])
    r1 = makePoints1D(1000.0, sp, slv, 5.0)
    assert r1.shape == (2, 3, 200) and np.allclose(r1[1, 1], np.arange(0, 1000, 5)) and np.all(r1[1, 2] == 0)
    with pytest.raises(ValueError):
        makePoints1D(10.0, sp.astype(np.float32), slv, 5.0)


# ---------------------------------------------------------------------------------------- RAiDER.interpolate
def test_interpolate_bit_exact_vs_reference_natives(gpu, golden):
    from raider_b200.interpolate import interpolate, interpolate_along_axis
    g = golden('interpolate')
    for nd in (1, 2, 3, 4):
        grids = [g[f'nd{nd}_g{d}'] for d in range(nd)]
        vals, pts = g[f'nd{nd}_vals'], g[f'nd{nd}_pts']
        assert np.array_equal(interpolate(grids, vals, pts, fill_value=np.nan), g[f'nd{nd}_fill'], equal_nan=True), nd
        assert np.array_equal(interpolate(grids, vals, pts), g[f'nd{nd}_clamp'], equal_nan=True), nd
        assert np.array_equal(interpolate(grids, vals, pts, assume_sorted=False, max_threads=1), g[f'nd{nd}_clamp'], equal_nan=True)
    assert np.array_equal(interpolate_along_axis(g['ax_x'], g['ax_y'], g['ax_new'], axis=2, fill_value=np.nan), g['ax_fill'], equal_nan=True)
    assert np.array_equal(interpolate_along_axis(g['ax_x'], g['ax_y'], g['ax_new'], axis=2), g['ax_clamp'], equal_nan=True)
    x1, y1, n1 = (np.ascontiguousarray(np.moveaxis(g[k], 2, 1)) for k in ('ax_x', 'ax_y', 'ax_new'))
    assert np.array_equal(interpolate_along_axis(x1, y1, n1, axis=1, fill_value=np.nan), g['ax1_fill'], equal_nan=True)


def test_interpolate_error_conventions(gpu):
    """module.cpp:36-63,309-347: TypeError for shape problems, RuntimeError for axis 0 with threads."""
    from raider_b200.interpolate import interpolate, interpolate_along_axis
    with pytest.raises(TypeError):
        interpolate(points=(np.zeros((10,)), np.zeros((5,))), values=np.zeros((1,)), interp_points=np.zeros((1,)))
    with pytest.raises(TypeError):
        interpolate_along_axis(np.array(0), np.array(0), np.array(0))
    with pytest.raises(TypeError):
        interpolate_along_axis(np.zeros(1), np.zeros(1), np.zeros((1, 1)))
    with pytest.raises(TypeError):
        interpolate_along_axis(np.zeros(1), np.zeros(2), np.zeros(1))
    with pytest.raises(TypeError):
        interpolate_along_axis(np.zeros(1), np.zeros(1), np.zeros(1), axis=1)
    with pytest.raises(TypeError):
        interpolate_along_axis(np.zeros((2, 2)), np.zeros((2, 2)), np.zeros((3, 2)))
    with pytest.raises(RuntimeError):
        interpolate_along_axis(np.zeros((2, 2)), np.zeros((2, 2)), np.zeros((2, 2)), axis=0)


def test_interpolate_reference_cases(gpu):
    """Numeric cases of test/test_interpolator.py (test_small, test_exact_points, *_out_of_bounds, *_fill_value, wrapper)."""
    from scipy.interpolate import RegularGridInterpolator as RGI
    from raider_b200.interpolate import interpolate, interpolate_along_axis
    from raider_b200.interpolator import RegularGridInterpolator as Interpolator, interp_along_axis, interpVector
    xs = np.array([1, 2, 3, 4, 5, 6]); ys = np.array([10, 9, 30, 10, 6, 1])
    ans = interpolate(points=(xs,), values=ys, interp_points=np.array([1.25, 2.9, 3.01, 5.7]).reshape(-1, 1))
    assert ans.shape == (4,) and np.allclose(ans, [9.75, 27.9, 29.8, 2.5], atol=1e-15)
    assert np.allclose(interpolate((xs,), ys, xs.reshape(-1, 1)), ys, atol=1e-15)
    assert interpolate((np.array([0, 1]),), np.array([0, 1]), np.array([[100]]), max_threads=1, assume_sorted=True) == np.array([100])
    assert np.all(np.isnan(interpolate((np.array([0, 1]),), np.array([0, 1]), np.array([[100]]), fill_value=np.nan)))
    g2 = (np.array([0, 1]), np.array([0, 1]))
    v2 = np.add.outer([0, 1], [0, 1])
    assert interpolate(g2, v2, np.array([[0.5, 0.5]])) == np.array([1])
    assert interpolate(g2, v2, np.array([[100, 100]])) == np.array([200])
    g4 = (np.array([0, 1]),) * 4
    v4 = np.add.outer(np.add.outer(v2, [0, 1]), [0, 1])
    assert interpolate(g4, v4, np.array([[0.5, 0.5, 0.5, 0.5]])) == np.array([2])
    assert interpolate(g4, v4, np.array([[100, 100, 100, 100]])) == np.array([400])
    # test_3d_cube_large-style: 2e5 points vs scipy
    f = lambda x, y, z: x ** 2 + 3 * y - z
    ax = np.linspace(0, 1000, 100)
    values = f(*np.meshgrid(ax, ax, ax, indexing='ij', sparse=True))
    n = 200_000
    pts = np.stack((np.linspace(10, 990, n), np.linspace(10, 890, n), np.linspace(10, 780, n)), axis=-1)
    assert np.allclose(interpolate((ax, ax, ax), values, pts, assume_sorted=True), RGI((ax, ax, ax), values)(pts), 1e-15)
    # wrapper (test_interpolate_wrapper): tuple-of-arrays and (N,3) forms, fill beyond the grid
    px, py, pz = np.linspace(10, 1090, 5), np.linspace(10, 890, 5), np.linspace(10, 890, 5)
    interp = Interpolator((ax, ax, ax), values, fill_value=np.nan)
    want = RGI((ax, ax, ax), values, bounds_error=False)(np.stack((px, py, pz), axis=-1))
    assert np.allclose(interp(np.stack((px, py, pz), axis=-1)), want, 1e-15, equal_nan=True)
    assert np.allclose(interp((px, py, pz)), want, 1e-15, equal_nan=True)
    # along-axis family (test_interp_along_axis*, test_interpVector)
    z2 = np.tile(np.arange(100)[..., np.newaxis], (5, 1, 5)).swapaxes(1, 2).astype(float)
    newz = np.tile(np.array([1.5, 9.9, 15, 23.278, 39.99, 50.1])[..., np.newaxis], (5, 1, 5)).swapaxes(1, 2)
    assert np.allclose(interp_along_axis(z2, newz, 0.3 * z2 - 12.75, axis=2), 0.3 * newz - 12.75)
    assert np.allclose(interpolate_along_axis(z2, 0.3 * z2 - 12.75, newz, axis=2), 0.3 * newz - 12.75)
    x1 = np.array([1, 2, 3, 4.0])
    assert np.allclose(interp_along_axis(x1, np.array([1.5, 3.1]), 2 * x1, axis=0), [3.0, 6.2])
    assert np.allclose(interp_along_axis(x1, np.array([0, 5.0]), 2 * x1, axis=0), [np.nan, np.nan], equal_nan=True)
    assert np.allclose(interpVector(np.array([0, 1, 2, 3, 4, 5, 0, 0.84147098, 0.90929743, 0.14112001, -0.7568025, -0.95892427,
                                              0.5, 1.5, 2.5, 3.5, 4.5]), 6),
                       [0.42073549, 0.87538421, 0.52520872, -0.30784124, -0.85786338])
    # threads edge case + large 3-D along-axis with scale (test_interp_along_axis_3d_large)
    scale = np.arange(1, 31).reshape((30, 1, 1))
    a2 = np.repeat(np.array([np.arange(100.0)]), 30, axis=0)
    xs3 = np.repeat(np.array([a2]), 30, axis=0) * scale
    p3 = np.repeat(np.array([np.array([np.linspace(0, 99, num=200)]).repeat(30, axis=0)]), 30, axis=0) * scale
    assert np.allclose(interpolate_along_axis(xs3, 2 * xs3, p3, axis=2, assume_sorted=True), 2 * p3)


# ---------------------------------------------------------------------------------------- K2
def test_sampler_bit_exact_vs_scipy(gpu, golden):
    """scipy RGI as delayFcns.py:55-56 configures it, incl. edges, OOB and NaN coordinates (Appendix A)."""
    from raider_b200 import _lib
    from raider_b200.delayFcns import getInterpolators
    g = golden('scipy_sample')
    ifW, ifH = getInterpolators({k: g[k[0] + 's'] if k in 'xyz' else g[k] for k in ('x', 'y', 'z', 'wet', 'hydro')})
    assert [a.size for a in ifW.grid] == [g['ys'].size, g['xs'].size, g['zs'].size]
    w, h = ifW(g['pts']), ifH(g['pts'])
    assert np.array_equal(w, g['out_wet'], equal_nan=True)
    assert np.array_equal(h, g['out_hydro'], equal_nan=True)
    assert np.isnan(w[4]) and np.isnan(w[5]) and np.isnan(w[6]) and not np.isnan(w[1])
    # shape handling like scipy: (..., 3) in -> (...) out
    assert ifW(g['pts'].reshape(30, 100, 3)).shape == (30, 100)
    # fp32 tier of the sampler (20 B/point): same cells, fp32 I/O
    w32, _ = ifW.cube.sample(g['pts'].astype(np.float32))
    ok = ~np.isnan(g['out_wet']) & ~np.isnan(w32)
    assert ok.sum() > 2500 and np.abs(w32[ok] - g['out_wet'][ok]).max() < 2e-2  # fp32 coordinates move the point by ~1e-3 m in z
    # C++ interval rules on the staged cube == oracle restatement of interpolate.cpp on the promoted values
    from oracle import interp as ointerp
    vals = g['wet'].transpose(1, 2, 0).astype(np.float64)
    for sem, fill in ((_lib.SEM_RAIDER_FILL, np.nan), (_lib.SEM_RAIDER_CLAMP, None)):
        got = ifW.cube.sample(g['pts'], semantics=sem)[0]
        want = ointerp.interpolate((g['ys'], g['xs'], g['zs']), vals, g['pts'], fill_value=fill)
        assert np.array_equal(got, want, equal_nan=True)


def test_sampler_descending_axes_and_layouts(gpu, golden):
    """Descending y (AOI grids, llreader.py:191) is flipped on staging like scipy does; (y,x,z) and (z,y,x) layouts agree."""
    from raider_b200 import _lib
    from raider_b200.engine import DeviceCube
    g = golden('scipy_sample')
    a = DeviceCube(g['ys'], g['xs'], g['zs'], g['wet'], g['hydro'], layout=_lib.LAYOUT_ZYX)
    b = DeviceCube(g['ys'][::-1].copy(), g['xs'], g['zs'], np.ascontiguousarray(g['wet'][:, ::-1]), np.ascontiguousarray(g['hydro'][:, ::-1]))
    c = DeviceCube(g['ys'], g['xs'], g['zs'], np.ascontiguousarray(g['wet'].transpose(1, 2, 0)),
                   np.ascontiguousarray(g['hydro'].transpose(1, 2, 0)), layout=_lib.LAYOUT_YXZ)
    ra = a.sample(g['pts'])
    for other in (b, c):
        ro = other.sample(g['pts'])
        assert np.array_equal(ra[0], ro[0], equal_nan=True) and np.array_equal(ra[1], ro[1], equal_nan=True)
    assert np.array_equal(b.grid[0], g['ys'])
    with pytest.raises(TypeError):
        DeviceCube(np.array([0.0, 1.0, 1.0]), g['xs'], g['zs'], g['wet'][:, :3], g['hydro'][:, :3])


def test_zenith_cube_matches_golden(gpu, golden):
    """_build_cube (delay.py:196-216) on the C1 cube: bit-exact vs scipy on wet_total/hydro_total."""
    from raider_b200 import synthetic as syn
    from raider_b200.delay import _build_cube
    from raider_b200.delayFcns import getInterpolators
    g = golden('raytrace')
    c1 = syn.config_c1()
    ifs = getInterpolators(c1['cube'], 'total')
    out = _build_cube(g['e_xpts'], g['e_ypts'], g['e_zpts'], 4326, 4326, list(ifs))
    assert out[0].shape == (5, 20, 20)
    assert np.array_equal(out[0], g['e_wet']) and np.array_equal(out[1], g['e_hydro'])


# ---------------------------------------------------------------------------------------- geodesy / build_ray
def test_geodesy_and_build_ray_vs_oracle(gpu, golden):
    from raider_b200.losreader import build_ray, getTopOfAtmosphere
    from raider_b200.utilFcns import ecef2lla, lla2ecef
    g = golden('geodesy')
    x, y, z = lla2ecef(g['lat'], g['lon'], g['h'])
    assert max(np.abs(x - g['x']).max(), np.abs(y - g['y']).max(), np.abs(z - g['z']).max()) < 5e-9  # ulp-level (sincos vs libm)
    lo, la, hh = ecef2lla(g['x'], g['y'], g['z'])
    assert np.abs(la - g['lat_back']).max() < 1e-13 and np.abs(hh - g['h_back']).max() < 5e-9
    dlon = np.abs(lo - g['lon_back'])
    assert dlon[np.abs(g['lat']) < 89.99].max() < 1e-13
    assert lla2ecef(0.0, 0.0, 0.0) == pytest.approx((6378137.0, 0.0, 0.0), abs=1e-9)   # test_delayFcns.py:86-99
    # Newton schedule: 10 iterations without factor, 3 with (losreader.py:720-733)
    assert np.abs(getTopOfAtmosphere(g['g0'], g['look'], 30000.0) - g['toa10']).max() < 1e-7
    assert np.abs(getTopOfAtmosphere(g['g0'], g['look'], 30000.0, factor=g['cosf']) - g['toa3']).max() < 1e-7
    lens, lows, highs = build_ray(g['zs'], 0.0, g['g0'], g['look'], g['zs'][-1] - 1)
    assert lens.shape == g['lens'].shape and lows.shape == g['lows'].shape
    assert np.abs(lens - g['lens']).max() < 1e-7 and np.abs(highs - g['highs']).max() < 1e-7 and np.abs(lows - g['lows']).max() < 1e-7
    assert build_ray(g['zs'], g['zs'][-1] - 0.5, g['g0'], g['look'], g['zs'][-1] - 1) == (None, None, None)  # losreader.py:832-833


# ---------------------------------------------------------------------------------------- K0 + K3
def _run_gpu(cfg, los, **kw):
    from raider_b200.delay import _build_cube_ray
    from raider_b200.delayFcns import getInterpolators
    ifs = getInterpolators(cfg['cube'])
    out = _build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], los, cfg.get('crs', 4326), 4326, list(ifs), MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                          MAX_TROPO_HEIGHT=cfg['zref'], **kw)
    return out, ifs[0].cube.last_info


def test_slant_fixed_incidence_golden_a(gpu, golden):
    """C2 shape: 30 deg incidence, NZ=37, 225 m segments; nParts bit-exact, delays within 1e-6 m of the oracle golden."""
    from raider_b200.losreader import Raytracing
    g = golden('raytrace')
    cfg = _c2_small(24, 0.08)
    out, info = _run_gpu(cfg, Raytracing(incidence=30.0, heading=-168.0))
    assert np.array_equal(info[0].nparts, g['a_nparts'])
    assert info[0].samples_per_ray == 296 and info[0].n_layers == 33 and info[0].reruns == 0
    assert np.abs(info[0].maxlen - g['a_maxlen']).max() < 1e-7
    assert np.abs(out[0] - g['a_wet']).max() < TOL_F64_M and np.abs(out[1] - g['a_hydro']).max() < TOL_F64_M
    assert np.abs(out[1] - g['a_hydro']).max() < 5e-10  # what the arithmetic really achieves (24 km span cubics: 1.5e-10 here)
    assert not np.isnan(out[0]).any()


def test_slant_ml145_two_heights_golden_b(gpu, golden):
    from raider_b200.losreader import Raytracing
    g = golden('raytrace')
    cfg = _c2_small(12, 0.15, table='ml145')
    cfg['zpts'] = np.array([0.0, 1500.0])
    out, info = _run_gpu(cfg, Raytracing(incidence=45.0, heading=12.0))
    assert np.array_equal(info[0].nparts, g['b_nparts']) and np.array_equal(info[1].nparts, g['b_nparts1'])
    assert out[0].shape == (2, 12, 12)
    assert np.abs(out[0] - g['b_wet']).max() < TOL_F64_M and np.abs(out[1] - g['b_hydro']).max() < TOL_F64_M


def test_slant_zenith_rays_golden_c(gpu, golden):
    from raider_b200.losreader import ZenithRaytracing
    g = golden('raytrace')
    cfg = _c2_small(10, 0.2)
    out, info = _run_gpu(cfg, ZenithRaytracing())
    assert np.array_equal(info[0].nparts, g['c_nparts'])
    assert np.abs(out[0] - g['c_wet']).max() < TOL_F64_M and np.abs(out[1] - g['c_hydro']).max() < TOL_F64_M


def test_slant_explicit_los_array_golden_d(gpu, golden):
    """Per-pixel look vectors handed over as an array (the contract of Raytracing.getLookVectors, losreader.py:219-255)."""
    from raider_b200.losreader import Raytracing
    g = golden('raytrace')
    cfg = _c2_small(16, 0.1)
    out, info = _run_gpu(cfg, Raytracing(look_vecs=g['d_los']))
    assert np.array_equal(info[0].nparts, g['d_nparts'])
    assert np.abs(out[0] - g['d_wet']).max() < TOL_F64_M and np.abs(out[1] - g['d_hydro']).max() < TOL_F64_M
    # a duck-typed third-party LOS object (only getLookVectors) takes the host geometry path and must agree
    class Duck:
        def getLookVectors(self, ht, llh, xyz, yy):
            assert xyz.shape == yy.shape + (3,)
            return g['d_los']
    out2, _ = _run_gpu(cfg, Duck())
    assert np.array_equal(out2[0], out[0]) and np.array_equal(out2[1], out[1])


def test_slant_hrrr_lcc_golden_f(gpu, golden):
    """HRRR-like cube: spherical LCC model CRS (models/hrrr.py:255-260), geographic query raster."""
    from raider_b200.crs import LambertConformalSphere
    from raider_b200.delay import _build_cube_ray
    from raider_b200.delayFcns import getInterpolators
    from raider_b200.losreader import Raytracing
    from raider_b200 import synthetic as syn
    g = golden('raytrace')
    lcc = LambertConformalSphere()
    assert np.allclose(lcc.params(), g['f_lcc'], rtol=1e-15)
    cube = {'x': g['f_xs'], 'y': g['f_ys'], 'z': syn.z_levels_table('hrrr57'), 'wet': g['f_wet_cube'], 'hydro': g['f_hydro_cube'], 'crs': lcc}
    ifs = getInterpolators(cube)
    out = _build_cube_ray(g['f_xpts'], g['f_ypts'], np.array([200.0]), Raytracing(incidence=35.0, heading=-12.0), lcc, 4326, list(ifs),
                          MAX_TROPO_HEIGHT=float(cube['z'][-1] - 1))
    assert np.array_equal(ifs[0].cube.last_info[0].nparts, g['f_nparts'])
    assert np.abs(out[0] - g['f_wet']).max() < TOL_F64_M and np.abs(out[1] - g['f_hydro']).max() < TOL_F64_M


def test_slant_vs_oracle_live_and_accumulate(gpu):
    """Same seeded inputs through the oracle here and now (not a stored vector), plus outputArrs accumulation (delay.py:245-248)."""
    from oracle import raytrace as rt
    from raider_b200.losreader import Raytracing
    cfg = _c2_small(20, 0.09, nz=50)
    cfg['zpts'] = np.array([0.0, 300.0, 2500.0])
    crs = rt.GeographicCRS()
    st = {}
    want = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], rt.FixedIncidenceLOS(38.0, 191.0), crs, crs,
                             list(rt.get_interpolators(cfg['cube'])), MAX_SEGMENT_LENGTH=400.0, MAX_TROPO_HEIGHT=cfg['zref'], stats=st)
    cfg['max_segment_length'] = 400.0
    out, info = _run_gpu(cfg, Raytracing(incidence=38.0, heading=191.0))
    for hh in range(3):
        assert np.array_equal(info[hh].nparts, st['nParts'][hh])
    assert np.abs(out[0] - want[0]).max() < TOL_F64_M and np.abs(out[1] - want[1]).max() < TOL_F64_M
    pre = [np.full_like(out[0], 1.0), np.full_like(out[1], -2.0)]
    ret, _ = _run_gpu(cfg, Raytracing(incidence=38.0, heading=191.0), outputArrs=pre)
    assert ret is None and np.allclose(pre[0], 1.0 + out[0], atol=1e-15) and np.allclose(pre[1], out[1] - 2.0, atol=1e-15)


def test_constant_refractivity_identity_on_device(gpu):
    """test/test_synthetic.py:217-274 on the GPU: delay * 1e6 == k * sum_k L_k, ray lengths recomputed independently with build_ray."""
    from raider_b200 import synthetic as syn
    from raider_b200.losreader import Raytracing, build_ray
    from raider_b200.utilFcns import lla2ecef
    cfg = _c2_small(32, 0.06, table='ml145')
    cube = syn.constant_cube(cfg['cube']['y'], cfg['cube']['x'], cfg['cube']['z'], 77.6, 71.6)
    los = Raytracing(incidence=41.0, heading=-168.0)
    out, info = _run_gpu(dict(cfg, cube=cube), los)
    xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
    xyz = np.stack(lla2ecef(yy, xx, np.zeros_like(yy)), -1)
    L = build_ray(cube['z'], 0.0, xyz, los.getLookVectors(0.0, [xx, yy, 0 * yy], xyz, yy), cfg['zref'])[0].sum(0)
    assert np.all(L > 1)
    for arr, k in ((out[0][0], np.float32(77.6)), (out[1][0], np.float32(71.6))):
        resid = (float(k) * L - arr * 1e6) / (float(k) * L)
        np.testing.assert_almost_equal(0, resid, decimal=9)   # the reference asks for 6 decimals


def test_whole_raster_predicates(gpu):
    """delay.py:276-277 (top slice skipped), :306-307 (all pixels below min(z) -> clamp), partial OOB -> NaN pixels, no error."""
    from oracle import raytrace as rt
    from raider_b200.losreader import Raytracing
    from raider_b200 import synthetic as syn
    cfg = _c2_small(8, 0.2)
    zs = cfg['cube']['z']
    # (1) default height_levels = model levels: the last level contributes nothing and is skipped, zeros stay
    cfg['zpts'] = np.array([zs[3], zs[-1]])
    out, info = _run_gpu(cfg, Raytracing(incidence=30.0, heading=-168.0))
    assert info[1].skipped and np.all(out[0][1] == 0) and np.all(out[0][0] > 0)
    # (2) ... but a non-final height without layers is the reference's latent TypeError
    cfg['zpts'] = np.array([zs[-1], zs[3]])
    with pytest.raises(TypeError):
        _run_gpu(cfg, Raytracing(incidence=30.0, heading=-168.0))
    # (3) ht == min(z): the first sample sits within ~1e-9 m of the cube floor; whatever the knife edge decides,
    #     the device must take the same branch as its own evaluation of the predicate and stay NaN-consistent
    cfg['zpts'] = np.array([zs[0]])
    out, info = _run_gpu(cfg, Raytracing(incidence=30.0, heading=-168.0))
    n_nan = int(np.isnan(out[0]).sum())
    assert (n_nan == 0) if info[0].clamp_low_first else (n_nan == info[0].oob_below)
    crs = rt.GeographicCRS()
    want = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], rt.FixedIncidenceLOS(30.0, -168.0), crs, crs,
                             list(rt.get_interpolators(cfg['cube'])), MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
    both = ~np.isnan(out[0]) & ~np.isnan(want[0])
    assert both.sum() > 0 and np.abs(out[0][both] - want[0][both]).max() < TOL_F64_M
    # (4) raster hanging over the edge of the cube: those pixels are NaN (delay.py:187-188 only logs), the rest are right
    xp = np.linspace(cfg['cube']['x'][-1] - 0.3, cfg['cube']['x'][-1] + 0.2, 9)
    cfg2 = dict(cfg, xpts=xp, zpts=np.array([0.0]))
    out, _ = _run_gpu(cfg2, Raytracing(incidence=30.0, heading=-168.0))
    want = rt.build_cube_ray(xp, cfg['ypts'], np.array([0.0]), rt.FixedIncidenceLOS(30.0, -168.0), crs, crs,
                             list(rt.get_interpolators(cfg['cube'])), MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
    assert np.array_equal(np.isnan(out[0]), np.isnan(want[0])) and np.isnan(out[0]).any() and not np.isnan(out[0]).all()
    ok = ~np.isnan(want[0])
    assert np.abs(out[0][ok] - want[0][ok]).max() < TOL_F64_M


def test_two_epoch_blend(gpu):
    """Temporal interpolation fused at staging (cli/raider.py:817-819) == blend cubes first, then trace."""
    from raider_b200 import synthetic as syn
    from raider_b200.delayFcns import getInterpolators
    from raider_b200.delay import _build_cube_ray
    from raider_b200.losreader import Raytracing
    cfg = _c2_small(10, 0.18)
    c0 = cfg['cube']
    c1 = syn.make_cube(c0['y'], c0['x'], c0['z'], seed=20200131)
    w0, w1 = 1 - 5 / 180, 5 / 180   # 12:05 between 12:00 and 15:00 (get_weights_time_interp, cli/raider.py:877-888)
    los = Raytracing(incidence=30.0, heading=-168.0)
    kw = dict(MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
    ifs = getInterpolators(c0)
    ifs[0].cube.blend(c1['wet'], c1['hydro'], w0, w1)
    fused = _build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], los, 4326, 4326, list(ifs), **kw)
    pre = _build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], los, 4326, 4326, list(getInterpolators(syn.blend_cubes(c0, c1, w0, w1))), **kw)
    assert np.abs(fused[0] - pre[0]).max() < 1e-7 and np.abs(fused[1] - pre[1]).max() < 1e-7  # fp32 (xarray) vs fp64 blend rounding


def test_tropo_delay_entry_points(gpu, tmp_path):
    """tropo_delay (delay.py:35-130): cube AOI -> Dataset; point AOI -> arrays; zenith, projected and ray-traced."""
    import datetime as dt
    from oracle import raytrace as rt
    from raider_b200 import synthetic as syn
    from raider_b200.cube_io import write_cube
    from raider_b200.delay import tropo_delay
    from raider_b200.llreader import BoundingBox, Points
    from raider_b200.losreader import Conventional, Raytracing, Zenith
    cfg = _c2_small(8, 0.2)
    path = write_cube(tmp_path / 'ERA5_synth.nc', cfg['cube'])
    t = dt.datetime(2020, 1, 30, 13, 52, 45)
    aoi = BoundingBox([33.2, 34.8, -118.8, -117.2], spacing=0.2)
    hts = [0.0, 500.0, 4000.0]
    crs = rt.GeographicCRS()
    # zenith cube
    ds, _ = tropo_delay(t, path, aoi, Zenith(), height_levels=hts)
    want = rt.build_cube(aoi.xpts, aoi.ypts, np.array(hts), crs, crs, list(rt.get_interpolators(cfg['cube'], 'total')))
    assert np.array_equal(np.asarray(ds['wet'].values), want[0]) and np.array_equal(np.asarray(ds['hydro'].values), want[1])
    assert ds['wet'].attrs['units'] == 'm' and 'zenith' in ds['wet'].attrs['description']
    # ray-traced cube, zref defaulting to top-of-model - 1 (delay.py:78-93)
    ds, _ = tropo_delay(t, path, aoi, Raytracing(incidence=30.0, heading=-168.0), height_levels=hts)
    want = rt.build_cube_ray(aoi.xpts, aoi.ypts, np.array(hts), rt.FixedIncidenceLOS(30.0, -168.0), crs, crs,
                             list(rt.get_interpolators(cfg['cube'])), MAX_TROPO_HEIGHT=cfg['cube']['z'].max() - 1)
    assert np.abs(np.asarray(ds['wet'].values) - want[0]).max() < TOL_F64_M
    assert np.abs(np.asarray(ds['hydro'].values) - want[1]).max() < TOL_F64_M
    # point mode: cube first, then scipy-style interpolation at (lat, lon, hgt), then projection (delay.py:98-128)
    rng = np.random.default_rng(7)
    lats, lons, hgts = rng.uniform(33.3, 34.7, 50), rng.uniform(-118.7, -117.3, 50), rng.uniform(10, 3000, 50)
    pts_aoi = Points(lats, lons, hgts)
    pts_aoi.xpts, pts_aoi.ypts = aoi.xpts, aoi.ypts
    from scipy.interpolate import RegularGridInterpolator as RGI
    wz, hz = tropo_delay(t, path, pts_aoi, Zenith(), height_levels=hts)
    wantz = rt.build_cube(aoi.xpts, aoi.ypts, np.array(hts), crs, crs, list(rt.get_interpolators(cfg['cube'], 'total')))
    ref_w = RGI((aoi.ypts, aoi.xpts, np.array(hts)), wantz[0].transpose(1, 2, 0), fill_value=np.nan, bounds_error=False)(np.stack([lats, lons, hgts], -1))
    assert np.abs(wz - ref_w).max() < 1e-12
    wp, hp = tropo_delay(t, path, pts_aoi, Conventional(incidence=35.0), height_levels=hts)
    assert np.allclose(wp, wz / np.cos(np.radians(35.0)), rtol=1e-14) and np.allclose(hp, hz / np.cos(np.radians(35.0)), rtol=1e-14)


def test_fp32_output_tier(gpu):
    """fp32 outputs (16 -> 8 B/ray of writes): within 1e-3 m of the fp64 oracle as north_star states for the fp32 tier."""
    import ctypes as C
    from raider_b200 import _lib
    from raider_b200.delayFcns import getInterpolators
    g = np.load(__import__('pathlib').Path(__file__).parent / 'golden' / 'raytrace.npz')
    cfg = _c2_small(24, 0.08)
    cube = getInterpolators(cfg['cube'])[0].cube
    enu = np.array([np.sin(np.radians(30)) * np.cos(np.radians(-168 + 90)), np.sin(np.radians(30)) * np.sin(np.radians(-168 + 90)), np.cos(np.radians(30))])
    maxlen, counts = cube.ray_layers(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], 24, 24, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
    w, h = np.empty((24, 24), np.float32), np.empty((24, 24), np.float32)
    nparts, oob = cube.ray_integrate(maxlen, cfg['max_segment_length'], False, w, h)
    assert np.array_equal(nparts, g['a_nparts'])
    assert np.abs(w - g['a_wet'][0]).max() < TOL_F32_M and np.abs(h - g['a_hydro'][0]).max() < TOL_F32_M


def test_full_size_properties(gpu):
    """BASELINE C2 at full size (2000 x 2000 rays) through size-independent properties: tiling invariance (sub-raster ==
    crop of the full raster when the global maxima are shared), linearity in the cube values, and the constant-N identity."""
    from raider_b200 import _lib, synthetic as syn
    from raider_b200.delayFcns import getInterpolators
    cfg = syn.config_c2()
    cube = getInterpolators(cfg['cube'])[0].cube
    enu = np.array([np.sin(np.radians(30)) * np.cos(np.radians(-78)), np.sin(np.radians(30)) * np.sin(np.radians(-78)), np.cos(np.radians(30))])
    ny = nx = 2000
    w, h = np.empty((ny, nx)), np.empty((ny, nx))
    info = cube.trace(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'], cfg['max_segment_length'], w, h)
    assert info.samples_per_ray in range(280, 320) and not np.isnan(w).any() and w.min() > 0.05 and h.min() > 2.0
    # tiling invariance: rows 700..899 alone, with the full raster's maxima injected
    ws, hs = np.empty((200, nx)), np.empty((200, nx))
    cube.ray_layers(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'][700:900], 200, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
    nparts, _ = cube.ray_integrate(info.maxlen, cfg['max_segment_length'], False, ws, hs)
    assert np.array_equal(nparts, info.nparts) and np.array_equal(ws, w[700:900]) and np.array_equal(hs, h[700:900])
    # linearity: doubling the cube doubles the delays (power-of-two scaling is exact in fp32 and fp64)
    c2 = dict(cfg['cube'], wet=cfg['cube']['wet'] * 2, hydro=cfg['cube']['hydro'] * 2)
    cube2 = getInterpolators(c2)[0].cube
    w2, h2 = np.empty((ny, nx)), np.empty((ny, nx))
    cube2.trace(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'], cfg['max_segment_length'], w2, h2)
    assert np.array_equal(w2, 2 * w) and np.array_equal(h2, 2 * h)
    # The oracle's OWN per-layer maxima of the full raster (nothing borrowed from the device): losreader.build_ray's restatement on
    # the border pixels + every 20th row / column; the maximum of the smooth length field sits on the border, the interior
    # sample checks that.  nParts must be the device's, bit for bit; the maxima agree to the 1e-8 m of K0's polynomial form.
    from oracle import raytrace as rt
    own_max = _oracle_full_raster_maxima(cfg, rt.FixedIncidenceLOS(30.0, -168.0), rt.GeographicCRS())
    own_np = np.ceil(own_max / cfg['max_segment_length']).astype(int) + 1
    assert np.array_equal(own_np, info.nparts) and np.abs(own_max - info.maxlen).max() < 5e-8
    # oracle on a 64 x 64 crop with ITS OWN full-raster maxima: the crop must reproduce the device within 1e-6 m
    crs = rt.GeographicCRS()
    sl = slice(968, 1032)
    want = rt.build_cube_ray(cfg['xpts'][sl], cfg['ypts'][sl], cfg['zpts'], rt.FixedIncidenceLOS(30.0, -168.0), crs, crs,
                             list(rt.get_interpolators(cfg['cube'])), MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                             MAX_TROPO_HEIGHT=cfg['zref'], layer_maxlen=[own_max])
    assert np.abs(w[sl, sl] - want[0][0]).max() < TOL_F64_M and np.abs(h[sl, sl] - want[1][0]).max() < TOL_F64_M
    assert np.abs(w[sl, sl] - want[0][0]).max() < 1e-9


def _oracle_full_raster_maxima(cfg, los, pts_crs, model_zs=None, step=20, ht=0.0):
    """Per-layer maxima of |P_hi - P_lo| over the full raster from the oracle alone: border + every `step`-th row / column."""
    from oracle import geodesy, raytrace as rt
    xp, yp = cfg['xpts'], cfg['ypts']
    bx = np.concatenate([xp, xp, np.full(yp.size, xp[0]), np.full(yp.size, xp[-1])])
    by = np.concatenate([np.full(xp.size, yp[0]), np.full(xp.size, yp[-1]), yp, yp])
    ix, iy = np.meshgrid(xp[::step], yp[::step])
    zs = cfg['cube']['z'] if model_zs is None else model_zs

    def maxima(xx, yy):
        xx, yy = xx.reshape(1, -1), yy.reshape(1, -1)
        llh = [xx, yy, np.full(yy.shape, ht)]
        xyz = np.stack(geodesy.lla2ecef(llh[1], llh[0], llh[2]), -1)
        return rt.build_ray(zs, ht, xyz, los.getLookVectors(ht, llh, xyz, yy), cfg['zref'])[0].max((1, 2))
    border, interior = maxima(bx, by), maxima(ix, iy)
    assert np.all(interior <= border), 'the longest ray of a layer is not on the border of the raster'
    return border


def test_full_size_crops_c5_and_c3(gpu):
    """64 x 64 crops out of FULL-SIZE C5 (2.4e8 rays, NZ = 72, row-tiled walk on one GPU) and C3 (8e7 rays over a 3 km Lambert cube,
    fixed incidence) against the oracle run with its own full-raster maxima: step counts bit-exact, delays within 1e-6 m."""
    from oracle import raytrace as rt
    from raider_b200 import _lib, synthetic as syn
    from raider_b200.engine import DeviceCube
    enu = np.array([np.sin(np.radians(30)) * np.cos(np.radians(-78)), np.sin(np.radians(30)) * np.sin(np.radians(-78)), np.cos(np.radians(30))])
    los = rt.FixedIncidenceLOS(30.0, -168.0)
    for name, cfg, model_crs in (('c5', syn.config_c5(), None), ('c3', syn.config_c3(table='hrrr57'), 'lcc')):
        ny, nx = cfg['ypts'].size, cfg['xpts'].size
        cube = DeviceCube.from_dict(cfg['cube'])
        import torch
        w = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
        h = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
        info = cube.trace(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'], cfg['max_segment_length'], w, h)
        torch.cuda.synchronize()
        own_max = _oracle_full_raster_maxima(cfg, los, rt.GeographicCRS(), step=200)
        own_np = np.ceil(own_max / cfg['max_segment_length']).astype(int) + 1
        assert np.array_equal(own_np, info.nparts), name
        assert np.abs(own_max - info.maxlen).max() < 5e-8, name
        mcrs = rt.LambertCRS(**cfg['crs'].args) if model_crs else rt.GeographicCRS()
        for r0, c0 in ((0, 0), (ny // 2 - 32, nx // 2 - 32), (ny - 64, nx - 64)):
            rs, cs = slice(r0, r0 + 64), slice(c0, c0 + 64)
            want = rt.build_cube_ray(cfg['xpts'][cs], cfg['ypts'][rs], cfg['zpts'], los, mcrs, rt.GeographicCRS(), list(rt.get_interpolators(cfg['cube'])),
                                     MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'], layer_maxlen=[own_max])
            gw, gh = w[rs, cs].cpu().numpy(), h[rs, cs].cpu().numpy()
            assert np.abs(gw - want[0][0]).max() < TOL_F64_M and np.abs(gh - want[1][0]).max() < TOL_F64_M, (name, r0, c0)
            assert np.abs(gh - want[1][0]).max() < 1e-9, (name, r0, c0)
        del cube, w, h


def test_table_division_is_ieee_exact(gpu):
    """The sampler never issues a DDIV: t = (x - g[i]) / (g[i+1] - g[i]) is n * RN(1/d) + Markstein corrections.  Two
    corrections must reproduce IEEE division bit for bit (that is what keeps K2 bit-identical to scipy)."""
    import ctypes as C
    m1, m2 = C.c_int64(-1), C.c_int64(-1)
    assert gpu.rdr_selftest_div(200_000_000, 12345, C.byref(m1), C.byref(m2), 0) == 0
    print(f'division self-test over 2e8 pairs: 1-step mismatches {m1.value}, 2-step mismatches {m2.value}')
    assert m1.value == 0 and m2.value == 0


def test_sampler_stream_variants_bit_exact_vs_scipy_live(gpu):
    """Every interval-lookup variant of the streaming sampler K2 against the installed scipy (the code the reference calls,
    delayFcns.py:55-56): exact-uniform axes (nodes rebuilt in registers), uniform axes (direct index + node table), irregular
    axes (bin table); node hits, the inclusive last node, out-of-bounds and NaN coordinates; host and device point streams."""
    import torch
    from scipy.interpolate import RegularGridInterpolator as RGI
    from raider_b200.engine import DeviceCube
    rng = np.random.default_rng(5)
    zs = np.concatenate([[-500.0, -120.0, 0.0, 35.5], 100.0 + 48000.0 * (np.arange(1, 30) / 29.0) ** 2])
    axes = {
        'exact-uniform 0.25 deg': (31.5 + 0.25 * np.arange(23), -121.25 + 0.3125 * np.arange(19)),
        'uniform, not exact': (np.linspace(31.5, 37.1, 23), -121.2 + 0.1 * np.arange(19)),
        'irregular': (np.sort(rng.uniform(31.5, 37.0, 23)), np.sort(rng.uniform(-121.0, -115.0, 19))),
    }
    for name, (ys, xs) in axes.items():
        wet = rng.normal(50.0, 20.0, (zs.size, ys.size, xs.size)).astype(np.float32)
        hydro = rng.normal(250.0, 30.0, wet.shape).astype(np.float32)
        n = 40000 + 77  # ragged tail after the 256-point tiles
        pts = np.stack([rng.uniform(ys[0] - 0.2, ys[-1] + 0.2, n), rng.uniform(xs[0] - 0.2, xs[-1] + 0.2, n),
                        rng.uniform(zs[0] - 50.0, zs[-1] + 50.0, n)], axis=-1)
        k = np.arange(0, 6000, 3)  # exact node hits on every axis (first and last node included)
        pts[k, 0] = ys[rng.integers(0, ys.size, k.size)]
        pts[k + 1, 1] = xs[rng.integers(0, xs.size, k.size)]
        pts[k + 2, 2] = zs[rng.integers(0, zs.size, k.size)]
        pts[6000] = (ys[-1], xs[-1], zs[-1])
        pts[6001] = (ys[0], xs[0], zs[0])
        pts[6002] = (np.nan, xs[3], zs[3])
        pts[6003] = (ys[3], np.nan, zs[3])
        pts[6004] = (ys[3], xs[3], np.nan)
        pts[6005] = (np.nextafter(ys[-1], np.inf), xs[3], zs[3])
        pts[6006] = (ys[3], np.nextafter(xs[0], -np.inf), zs[3])
        want_w = RGI((ys, xs, zs), wet.transpose(1, 2, 0), method='linear', bounds_error=False, fill_value=np.nan)(pts)
        want_h = RGI((ys, xs, zs), hydro.transpose(1, 2, 0), method='linear', bounds_error=False, fill_value=np.nan)(pts)
        cube = DeviceCube(ys, xs, zs, wet, hydro)
        got_w, got_h = cube.sample(pts)  # host stream (staged), fp64
        assert np.array_equal(got_w, want_w, equal_nan=True), name
        assert np.array_equal(got_h, want_h, equal_nan=True), name
        assert np.isnan(want_w).sum() > 100 and np.isfinite(want_w).sum() > 20000
        dw, dh = cube.sample(torch.from_numpy(pts).cuda())  # device stream
        torch.cuda.synchronize()
        assert np.array_equal(dw.cpu().numpy(), want_w, equal_nan=True), name
        assert np.array_equal(dh.cpu().numpy(), want_h, equal_nan=True), name
        # fp32 tier (k_sample_stream_f32: fp32 coordinates, fp32 arithmetic, fp32 values) against scipy evaluated at the very same
        # fp32 points: the NaN pattern (closed-box rule, NaN coordinates) is exact, values agree to fp32 rounding of t and the lerps
        pts32 = pts.astype(np.float32)
        p64 = pts32.astype(np.float64)
        ref_w = RGI((ys, xs, zs), wet.transpose(1, 2, 0), method='linear', bounds_error=False, fill_value=np.nan)(p64)
        ref_h = RGI((ys, xs, zs), hydro.transpose(1, 2, 0), method='linear', bounds_error=False, fill_value=np.nan)(p64)
        dev32 = cube.sample(torch.from_numpy(pts32).cuda())   # device stream: asynchronous on the handle's stream
        torch.cuda.synchronize()
        for got_w32, got_h32 in (cube.sample(pts32), tuple(t.cpu().numpy() for t in dev32)):
            assert got_w32.dtype == np.float32
            assert np.array_equal(np.isnan(got_w32), np.isnan(ref_w)), name
            ok = ~np.isnan(ref_w)
            assert np.abs(got_w32[ok] - ref_w[ok]).max() < 2e-5 * 150.0, (name, float(np.abs(got_w32[ok] - ref_w[ok]).max()))
            assert np.abs(got_h32[ok] - ref_h[ok]).max() < 2e-5 * 400.0, (name, float(np.abs(got_h32[ok] - ref_h[ok]).max()))


# ---------------------------------------------------------------------------------------- K6: orbit look vectors
def _s1_orbit():
    import datetime as dt
    from conftest import GOLDEN
    from raider_b200.losreader import get_orbit
    return get_orbit(str(GOLDEN / 'orbit_S1_sv.txt'), dt.datetime(2018, 11, 12, 23, 0, 30), 600)


def test_orbit_look_vectors_vs_oracle(gpu):
    """Device zero-Doppler solve (geo2rdr + Hermite orbit interpolation, losreader.py:219-255) against the oracle's
    restatement: look vectors to 1e-12, slant range to 1e-6 m, NaN pattern identical (targets outside the 70 s orbit span)."""
    from oracle import orbit as ob
    from raider_b200.losreader import get_radar_pos, orbit_look_vectors, state_to_los
    o = _s1_orbit()
    oo = ob.Orbit(o.time, o.position, o.velocity)
    rng = np.random.default_rng(3)
    lat, lon, hgt = rng.uniform(13.3, 16.6, 300), rng.uniform(100.3, 105.0, 300), rng.uniform(-100.0, 4000.0, 300)
    los, sr, az = orbit_look_vectors(o, lat, lon, hgt)
    want = ob.look_vectors_points(lat, lon, hgt, oo)
    assert np.array_equal(np.isnan(los), np.isnan(want)) and 20 < np.isnan(want[:, 0]).sum() < 200
    ok = ~np.isnan(want[:, 0])
    # isce3's Newton drops the acceleration term, so it converges linearly (ratio ~ slant range / orbit radius) and stops, by
    # the 1e-7 m rule on the slant range, ~4e-6 s short of the zero-Doppler time.  Two correct implementations can therefore
    # differ by one iteration on a rounding knife edge: <= 4e-6 s * 7.6 km/s / 800 km = 4e-8 in the unit vector.  Everywhere else
    # they agree to rounding.
    err = np.abs(los[ok] - want[ok]).max(axis=-1)
    assert err.max() < 2e-7 and np.mean(err < 1e-12) > 0.9
    from oracle import geodesy
    xyz = np.stack(geodesy.lla2ecef(lat, lon, hgt), -1)
    for i in np.flatnonzero(ok)[:40]:
        a, s = ob.geo2rdr(xyz[i], oo)
        assert abs(a - az[i]) < 1e-5 and abs(s - sr[i]) < 1e-6
    # Conventional's orbit branch: cos(look angle) (losreader.py:558-606)
    svs = np.concatenate([o.time[:, None], o.position, o.velocity], axis=1)
    f = state_to_los(svs, [lat[ok], lon[ok], hgt[ok]])
    up = geodesy.getZenithLookVecs(lat[ok], lon[ok], hgt[ok])
    assert np.abs(f - np.sum(want[ok] * up, -1)).max() < 2e-7
    ang, sr2 = get_radar_pos(np.stack([lat[ok], lon[ok], hgt[ok]], -1), o)
    assert np.abs(np.cos(np.radians(ang)) - f).max() < 1e-12 and np.abs(sr2 - sr[ok]).max() == 0.0


def test_slant_orbit_los_on_device_vs_oracle(gpu):
    """Raytracing(orbit) end to end: LOS solved on the device (RDR_LOS_ORBIT, never materialised on the host), K0, K3 against
    the oracle's loop with its own geo2rdr; the host-geometry route (getLookVectors -> array upload) gives the same map."""
    import datetime as dt
    from conftest import GOLDEN
    from oracle import orbit as ob, raytrace as rt
    from raider_b200 import synthetic as syn
    from raider_b200.losreader import Raytracing
    n = 10
    xpts, ypts = syn.raster(15.4, 102.5, n, n, 0.15)
    xs, ys = syn.cube_axes_around(xpts, ypts)
    zs = syn.z_levels(37)
    cfg = {'cube': syn.make_cube(ys, xs, zs, totals=False), 'xpts': xpts, 'ypts': ypts, 'zpts': np.array([0.0, 800.0]),
           'zref': 15000.0, 'max_segment_length': 1000.0}
    los = Raytracing(filename=str(GOLDEN / 'orbit_S1_sv.txt'), time=dt.datetime(2018, 11, 12, 23, 0, 30))
    assert los.getSensorDirection() == 'desc' and los.getLookDirection() == 'right'
    out, info = _run_gpu(cfg, los)
    o = los._orbit
    crs = rt.GeographicCRS()
    st = {}
    want = rt.build_cube_ray(xpts, ypts, cfg['zpts'], ob.OrbitLOS(ob.Orbit(o.time, o.position, o.velocity)), crs, crs,
                             list(rt.get_interpolators(cfg['cube'])), MAX_SEGMENT_LENGTH=1000.0, MAX_TROPO_HEIGHT=15000.0, stats=st)
    for hh in range(2):
        assert np.array_equal(info[hh].nparts, st['nParts'][hh])
    assert not np.isnan(want[0]).any()
    assert np.abs(out[0] - want[0]).max() < TOL_F64_M and np.abs(out[1] - want[1]).max() < TOL_F64_M
    assert np.median(np.abs(out[1] - want[1])) < 1e-10  # knife-edge pixels of the LOS solve may reach ~1e-7 m (see K6 test)

    class HostOnly:  # a third-party LOS object: only getLookVectors, evaluated on the host and uploaded
        def __init__(self, inner):
            self.inner = inner

        def getLookVectors(self, ht, llh, xyz, yy):
            return self.inner.getLookVectors(ht, llh, xyz, yy)

        def is_Zenith(self):
            return False

        def is_Projected(self):
            return False

        def ray_trace(self):
            return True
    out2, _ = _run_gpu(cfg, HostOnly(los))
    assert np.abs(out2[0] - out[0]).max() < 1e-12 and np.abs(out2[1] - out[1]).max() < 1e-12
    # a raster that leaves the orbit's coverage entirely -> the reference's ValueError (delay.py:279-280)
    far = dict(cfg)
    far['xpts'], far['ypts'] = syn.raster(15.4, 60.0, 4, 4, 0.1)
    fx, fy = syn.cube_axes_around(far['xpts'], far['ypts'])
    far['cube'] = syn.make_cube(fy, fx, zs, totals=False)
    with pytest.raises(ValueError, match='geo2rdr did not converge'):
        _run_gpu(far, los)


# ---------------------------------------------------------------------------------------- K5: station (point) mode, C4
def _oracle_stations(cube, lat, lon, hgt, vecs, zref, seg):
    """Each station as the reference would do it: a 1 x 1 raster at the station's height through _build_cube_ray."""
    from oracle import raytrace as rt
    crs = rt.GeographicCRS()
    ifs = list(rt.get_interpolators(cube))
    wet, hydro, ns = np.zeros(lat.size), np.zeros(lat.size), np.zeros(lat.size, dtype=np.int64)
    for i in range(lat.size):
        st = {}
        out = rt.build_cube_ray(lon[i:i + 1], lat[i:i + 1], hgt[i:i + 1], rt.ArrayLOS(vecs[i][None, None, :]), crs, crs, ifs,
                                MAX_SEGMENT_LENGTH=seg, MAX_TROPO_HEIGHT=zref, stats=st)
        wet[i], hydro[i], ns[i] = out[0][0, 0, 0], out[1][0, 0, 0], int(st['nParts'][0].sum())
    return wet, hydro, ns


def test_station_mode_c4_vs_oracle(gpu):
    """BASELINE C4 at oracle-feasible size: per-station height / incidence / heading, two weather epochs blended at staging
    (cli/raider.py:817-819), one warp per ray.  Per-ray sample counts bit-exact, delays within 1e-6 m."""
    from oracle import geodesy
    from raider_b200 import synthetic as syn
    from raider_b200.delay import slant_delay_points
    from raider_b200.delayFcns import getInterpolators
    from raider_b200.losreader import Raytracing
    cfg = syn.config_c4(n=160)
    lat, lon, hgt = cfg['lat'], cfg['lon'], cfg['hgt']
    los = Raytracing(incidence=cfg['incidence'], heading=cfg['heading'])
    wet, hydro = slant_delay_points(cfg['cube0'], lat, lon, hgt, los, zref=cfg['zref'], MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                                    second_epoch=cfg['cube1'], weights=cfg['weights'])
    blended = syn.blend_cubes(cfg['cube0'], cfg['cube1'], *cfg['weights'])
    enu = geodesy.inc_hd_to_enu(cfg['incidence'], cfg['heading'])
    vecs = geodesy.enu2ecef(enu[:, 0], enu[:, 1], enu[:, 2], lat, lon, hgt)
    w_ref, h_ref, ns_ref = _oracle_stations(blended, lat, lon, hgt, vecs, cfg['zref'], cfg['max_segment_length'])
    assert not np.isnan(w_ref).any()
    assert np.abs(wet - w_ref).max() < TOL_F64_M and np.abs(hydro - h_ref).max() < TOL_F64_M
    assert np.abs(hydro - h_ref).max() < 1e-10
    # the same stations through explicit ECEF vectors and with the cube pair already staged; integer contract: samples per ray
    ifs = getInterpolators(blended)
    w2, h2 = slant_delay_points(list(ifs), lat, lon, hgt, np.ascontiguousarray(vecs), zref=cfg['zref'], MAX_SEGMENT_LENGTH=cfg['max_segment_length'])
    assert np.abs(w2 - wet).max() < 1e-12 and np.abs(h2 - hydro).max() < 1e-12
    assert np.array_equal(ifs[0].cube.last_station_samples, ns_ref)


def test_station_mode_edges(gpu):
    """Stations above zref / outside the cube / below the first model level, zenith rays equal the raster path."""
    from raider_b200 import _lib, synthetic as syn
    from raider_b200.delay import _build_cube_ray, slant_delay_points
    from raider_b200.delayFcns import getInterpolators
    from raider_b200.losreader import ZenithRaytracing
    cfg = syn.config_c2(n=6)
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 6, 6, 0.2)
    ifs = getInterpolators(cfg['cube'])
    xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
    ras = _build_cube_ray(cfg['xpts'], cfg['ypts'], np.array([250.0]), ZenithRaytracing(), 4326, 4326, list(ifs), MAX_TROPO_HEIGHT=15000.0)
    w, h = slant_delay_points(list(ifs), yy.ravel(), xx.ravel(), np.full(xx.size, 250.0), ZenithRaytracing(), zref=15000.0)
    # a vertical ray has the same length for every pixel up to ~1e-9 m, so global and per-ray nParts coincide
    assert np.abs(w - ras[0][0].ravel()).max() < 1e-9 and np.abs(h - ras[1][0].ravel()).max() < 1e-9
    lat = np.array([34.0, 34.0, 80.0, 34.0])
    lon = np.array([-118.0, -118.0, -118.0, -118.0])
    hgt = np.array([20000.0, -900.0, 100.0, 14999.5])  # above zref, below the first level (clamped like delay.py:306-307), outside, < 1 m
    w, h = slant_delay_points(list(ifs), lat, lon, hgt, ZenithRaytracing(), zref=15000.0)
    assert w[0] == 0.0 and h[0] == 0.0 and w[3] == 0.0       # no contributing layer: zeros stay (delay.py:276-277)
    assert np.isfinite(w[1]) and w[1] > 0 and np.isnan(w[2])
    n0 = ifs[0].cube.last_station_samples
    assert n0[0] == 0 and n0[3] == 0 and n0[1] > 10


# ---------------------------------------------------------------------------------------- C3 / C5 at oracle-feasible sizes, tiling
def test_c3_lcc_cube_with_orbit_los_vs_oracle(gpu):
    """BASELINE C3 shape: HRRR-like spherical-LCC cube (3 km, NZ = 50), geographic raster, per-pixel LOS from a synthetic
    circular orbit solved on the device; against the oracle's loop (its own LCC forward + geo2rdr)."""
    from oracle import orbit as ob, raytrace as rt
    from raider_b200 import synthetic as syn
    from raider_b200.delay import _build_cube_ray
    from raider_b200.delayFcns import getInterpolators
    from raider_b200.losreader import Orbit, Raytracing
    cfg = syn.config_c3(ny=9, nx=11)
    rows = cfg['orbit_rows']
    los = Raytracing(filename=Orbit(rows[:, 0], rows[:, 1:4], rows[:, 4:7]))
    ifs = getInterpolators(cfg['cube'])
    out = _build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], los, cfg['crs'], 4326, list(ifs), MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                          MAX_TROPO_HEIGHT=cfg['zref'])
    lcc = rt.LambertCRS()
    assert np.allclose(lcc.params(), cfg['crs'].params(), rtol=1e-15)
    st = {}
    cube = {k: v for k, v in cfg['cube'].items() if k != 'crs'}
    want = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], ob.OrbitLOS(ob.Orbit(rows[:, 0], rows[:, 1:4], rows[:, 4:7])), lcc,
                             rt.GeographicCRS(), list(rt.get_interpolators(cube)), MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                             MAX_TROPO_HEIGHT=cfg['zref'], stats=st)
    assert np.array_equal(ifs[0].cube.last_info[0].nparts, st['nParts'][0])
    assert not np.isnan(want[0]).any()
    assert np.abs(out[0] - want[0]).max() < TOL_F64_M and np.abs(out[1] - want[1]).max() < TOL_F64_M


def test_row_tiling_is_bit_identical(gpu, monkeypatch):
    """HBM budget of the t-buffer: a raster walked in row tiles (global maxima first, then K0 + K3 per tile) gives the same
    bits as the untiled raster -- the mechanism C5 relies on when 2.4e8 rays x 73 layers do not fit one GPU."""
    from raider_b200 import synthetic as syn
    from raider_b200.losreader import Raytracing
    cfg = syn.config_c5(ny=61, nx=40)
    los = Raytracing(incidence=30.0, heading=-168.0)
    whole, info = _run_gpu(cfg, los)
    assert info[0].tiles == 1 and info[0].n_layers > 60
    monkeypatch.setenv('RAIDER_B200_T_BUDGET_GB', repr(8 * 72 * 40 * 7.5 / 2**30))  # 7 rows per tile -> 9 tiles
    tiled, info_t = _run_gpu(cfg, los)
    assert info_t[0].tiles == 9 and np.array_equal(info_t[0].nparts, info[0].nparts)
    assert np.array_equal(whole[0], tiled[0]) and np.array_equal(whole[1], tiled[1])
    from oracle import raytrace as rt
    crs = rt.GeographicCRS()
    want = rt.build_cube_ray(cfg['xpts'], cfg['ypts'][:6], cfg['zpts'], rt.FixedIncidenceLOS(30.0, -168.0), crs, crs,
                             list(rt.get_interpolators(cfg['cube'])), MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                             MAX_TROPO_HEIGHT=cfg['zref'], layer_maxlen=[info[0].maxlen])
    assert np.abs(tiled[0][0, :6] - want[0][0]).max() < TOL_F64_M and np.abs(tiled[1][0, :6] - want[1][0]).max() < TOL_F64_M


def test_k3_integrator_forms_agree(gpu, monkeypatch):
    """The three forms of K3 -- polynomial (default: per-span cubics of the cube coordinates through exact nodes), fast (Bowring
    per sample) and general (PROJ-form arithmetic per sample) -- against the oracle and against each other, on the geometries
    that stress the polynomial: steep incidence (long spans), high latitude, thick top layers, short spans, cached / uncached
    cell records, a Lambert cube, rays that leave the cube (handed to the PROJ-form path with their NaN pattern)."""
    from oracle import raytrace as rt
    from raider_b200 import synthetic as syn
    from raider_b200.losreader import Raytracing
    crs = rt.GeographicCRS()

    def forms(cfg, los, settings):
        res = {}
        for name, env in settings.items():
            for k in ('RDR_K3_MODE', 'RDR_K3_SPAN', 'RDR_K3_CACHE', 'RDR_K3_MINB', 'RDR_K0_MODE'):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            res[name] = _run_gpu(cfg, los)
        for k in ('RDR_K3_MODE', 'RDR_K3_SPAN', 'RDR_K3_CACHE', 'RDR_K3_MINB', 'RDR_K0_MODE'):
            monkeypatch.delenv(k, raising=False)
        return res

    # (+ K0 with Bowring heights at every Newton iterate instead of the span cubics of h(t): 'k0_exact')
    settings = {'poly': {}, 'k0_exact': {'RDR_K0_MODE': 'exact'}, 'poly_nocache': {'RDR_K3_CACHE': '0'}, 'poly_cache_m4': {'RDR_K3_CACHE': '1', 'RDR_K3_MINB': '4'},
                'poly_span2k': {'RDR_K3_SPAN': '2000'}, 'poly_span20k': {'RDR_K3_SPAN': '20000'}, 'fast': {'RDR_K3_MODE': 'fast'},
                'general': {'RDR_K3_MODE': 'general'}}
    cases = [(_c2_small(24, 0.05), 30.0, -168.0, 225.0), (_c2_small(24, 0.05), 62.0, 77.0, 500.0), (_c2_small(16, 0.05, table='ml145'), 45.0, 10.0, 1000.0)]
    hi = syn.config_c2(n=16)   # 72 N: meridians converge, the cube cells are 9 km wide in x
    hi['xpts'], hi['ypts'] = syn.raster(72.0, 25.0, 16, 16, 0.05)
    xs, ys = syn.cube_axes_around(hi['xpts'], hi['ypts'])
    hi['cube'] = syn.make_cube(ys, xs, syn.z_levels(37), totals=False)
    cases.append((hi, 40.0, -100.0, 225.0))
    for cfg, inc, head, seg in cases:
        cfg = dict(cfg, max_segment_length=seg)
        st = {}
        want = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], rt.FixedIncidenceLOS(inc, head), crs, crs, list(rt.get_interpolators(cfg['cube'])),
                                 MAX_SEGMENT_LENGTH=seg, MAX_TROPO_HEIGHT=cfg['zref'], stats=st)
        assert not np.isnan(want[0]).any()
        res = forms(cfg, Raytracing(incidence=inc, heading=head), settings)
        assert np.abs(res['poly'][1][0].maxlen - res['k0_exact'][1][0].maxlen).max() < 1e-7   # layer lengths: cubic vs Bowring iterates
        for name, (out, info) in res.items():
            assert np.array_equal(info[0].nparts, st['nParts'][0]), name
            for f in (0, 1):
                assert np.abs(out[f] - want[f]).max() < TOL_F64_M, (name, inc)
                # the forms differ from each other by rounding only: 1e-9 m is 1000x inside the contract
                assert np.abs(out[f] - res['general'][0][f]).max() < 1e-9, (name, inc, float(np.abs(out[f] - res['general'][0][f]).max()))
    # rays leaving the cube: the polynomial integrator hands them over; NaN pattern and values as the PROJ-form path alone
    m = 40
    xp, yp = syn.raster(34.0, -118.0, m, m, 0.004)
    xs, ys = syn.cube_axes_around(xp, yp, pad_deg=0.0)
    edge = {'cube': syn.make_cube(ys, xs, syn.z_levels(37), totals=False), 'xpts': xp, 'ypts': yp, 'zpts': np.array([0.0]),
            'zref': float(syn.z_levels(37)[-1] - 1), 'max_segment_length': 225.0}
    res = forms(edge, Raytracing(incidence=30.0, heading=-168.0), {'poly': {}, 'general': {'RDR_K3_MODE': 'general'}})
    a, b = res['poly'][0], res['general'][0]
    assert np.isnan(b[0]).any() and not np.isnan(b[0]).all()
    assert np.array_equal(np.isnan(a[0]), np.isnan(b[0])) and np.nanmax(np.abs(a[0] - b[0])) < 1e-9 and np.nanmax(np.abs(a[1] - b[1])) < 1e-9
    # Lambert cube (C3 shape): polynomial nodes through the PROJ-form inverse + Lambert forward vs every sample through them
    c3 = syn.config_c3(ny=12, nx=14)
    res = forms(c3, Raytracing(incidence=37.0, heading=-168.0), {'poly': {}, 'poly_nocache': {'RDR_K3_CACHE': '0'}, 'general': {'RDR_K3_MODE': 'general'}})
    for name in ('poly', 'poly_nocache'):
        for f in (0, 1):
            assert np.abs(res[name][0][f] - res['general'][0][f]).max() < 1e-9, name


def test_k0_very_oblique_rays_take_the_exact_form(gpu, monkeypatch):
    """80 deg incidence through the 145-level table: the ray is longer than K0's span table (384 km), the warp redoes its layers
    with Bowring heights; layer maxima / nParts must not depend on which form ran, and match the oracle.  The rays leave the
    cube on the way up (NaN delays, as in the reference: delay.py:187-188)."""
    from oracle import raytrace as rt
    from raider_b200.losreader import Raytracing
    cfg = _c2_small(8, 0.05, table='ml145')
    crs = rt.GeographicCRS()
    st = {}
    want = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], rt.FixedIncidenceLOS(80.0, -168.0), crs, crs, list(rt.get_interpolators(cfg['cube'])),
                             MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'], stats=st)
    out, info = _run_gpu(cfg, Raytracing(incidence=80.0, heading=-168.0))
    monkeypatch.setenv('RDR_K0_MODE', 'exact')
    out_e, info_e = _run_gpu(cfg, Raytracing(incidence=80.0, heading=-168.0))
    assert np.array_equal(info[0].maxlen, info_e[0].maxlen)          # the very same code ran
    assert np.array_equal(info[0].nparts, st['nParts'][0])
    assert np.array_equal(np.isnan(out[0]), np.isnan(want[0])) and np.isnan(want[0]).all()


def test_nan_nodes_poison_like_the_reference(gpu):
    """A NaN node of the cube poisons every sample whose cell touches it, zero weight or not (0 * NaN = NaN in
    scipy and in interpolate.cpp:165-174; cubes are only *warned* about, delayFcns.py:43-44): rays through such a cell come out
    NaN, all others are untouched -- through the closed-form layer sums, the per-sample loops and the PROJ-form integrator alike."""
    from oracle import raytrace as rt
    from raider_b200.losreader import Raytracing
    cfg = _c2_small(20, 0.09)
    cube = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in cfg['cube'].items()}
    ny, nx = cube['y'].size, cube['x'].size
    cube['wet'][20, ny // 2, nx // 2] = np.nan        # mid troposphere (thick layer: quadrature), one column
    cube['hydro'][3, ny // 2 - 1, nx // 2] = np.nan   # near the ground (thin layers: per-sample path), the column south of it
    cfg = dict(cfg, cube=cube)
    crs = rt.GeographicCRS()
    want = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], rt.FixedIncidenceLOS(30.0, -168.0), crs, crs, list(rt.get_interpolators(cube)),
                             MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
    assert 0 < np.isnan(want[0]).sum() < want[0].size and 0 < np.isnan(want[1]).sum() < want[1].size
    for mode in ('poly', 'general'):
        import os
        os.environ['RDR_K3_MODE'] = mode
        try:
            out, _ = _run_gpu(cfg, Raytracing(incidence=30.0, heading=-168.0))
        finally:
            os.environ.pop('RDR_K3_MODE', None)
        for f in (0, 1):
            assert np.array_equal(np.isnan(out[f]), np.isnan(want[f])), (mode, f)
            ok = ~np.isnan(want[f])
            assert np.abs(out[f][ok] - want[f][ok]).max() < TOL_F64_M, (mode, f)


def test_peer_outputs_mirror_the_maps(gpu):
    """rdr_set_peer_outputs (the all-gather fused into K3, raider_b200.dist.SymmetricMaps): every destination receives the same
    bits as the primary output -- here the "peers" are two more buffers on the same GPU, at a row offset inside larger maps, for
    the polynomial integrator, the PROJ-form list pass (rays leaving the cube) and the general integrator; += calls ignore them."""
    import torch
    from raider_b200 import _lib, synthetic as syn
    from raider_b200.engine import DeviceCube
    from raider_b200.losreader import inc_hd_to_enu
    m = 40
    xp, yp = syn.raster(34.0, -118.0, m, m, 0.004)
    xs, ys = syn.cube_axes_around(xp, yp, pad_deg=0.0)          # no padding: slanted rays leave the cube -> list pass runs
    zs = syn.z_levels(37)
    cube = DeviceCube.from_dict(syn.make_cube(ys, xs, zs, totals=False), device=0)
    enu = np.ascontiguousarray(inc_hd_to_enu(np.float64(30.0), np.float64(-168.0)))
    maxlen, counts = cube.ray_layers(_lib.GEOM_GRID, xp, yp, m, m, _lib.LOS_ENU_CONST, enu, 0.0, float(zs[-1] - 1))
    for mode in ('poly', 'general'):
        import os
        os.environ['RDR_K3_MODE'] = mode
        try:
            ow = torch.full((m, m), -1.0, dtype=torch.float64, device='cuda')
            oh = torch.full((m, m), -1.0, dtype=torch.float64, device='cuda')
            big = [torch.full((2, 3 * m, m), -7.0, dtype=torch.float64, device='cuda') for _ in range(2)]   # two "peer" map pairs
            row0 = m  # this rank's block starts at row m of the 3m-row maps
            peers = ([b[0, row0:].data_ptr() for b in big], [b[1, row0:].data_ptr() for b in big])
            cube.ray_integrate(maxlen, 225.0, False, ow, oh, peers=peers)
            torch.cuda.synchronize()
            assert cube.h.last_fix_count != 0 if mode == 'poly' else True
            for b in big:
                assert torch.equal(torch.nan_to_num(b[0, row0:row0 + m], nan=123.0), torch.nan_to_num(ow, nan=123.0))
                assert torch.equal(torch.nan_to_num(b[1, row0:row0 + m], nan=123.0), torch.nan_to_num(oh, nan=123.0))
                assert bool((b[:, :row0] == -7.0).all()) and bool((b[:, row0 + m:] == -7.0).all())     # nothing outside the block
            assert bool(torch.isnan(ow).any()) and bool(torch.isfinite(ow).any())
            # the list is consumed by the call: a later plain call writes no peer
            for b in big:
                b.fill_(-7.0)
            cube.ray_integrate(maxlen, 225.0, False, ow, oh)
            torch.cuda.synchronize()
            assert all(bool((b == -7.0).all()) for b in big)
        finally:
            os.environ.pop('RDR_K3_MODE', None)


# ---------------------------------------------------------------------------------------- K7: weather-model processing (f4)
def _native_columns(ny=6, nx=7, nl=40, seed=4):
    rng = np.random.default_rng(seed)
    base = np.concatenate([[-60.0, 30.0], 150.0 + 42000.0 * (np.arange(1, nl - 1) / (nl - 2)) ** 2])
    zs = base[None, None, :] * rng.normal(1.0, 0.01, (ny, nx, 1)) + rng.uniform(-40.0, 400.0, (ny, nx, 1))  # terrain-following columns
    t = 288.15 - 6.5e-3 * np.clip(zs, None, 11000.0) + rng.normal(0, 0.5, zs.shape)
    p = 101325.0 * np.exp(-zs / 8000.0) * (1 + rng.normal(0, 1e-3, zs.shape))
    q = 8e-3 * np.exp(-zs / 2500.0) * (1 + 0.2 * rng.random(zs.shape))
    return zs, p, t, q


def test_weather_processing_vs_oracle(gpu):
    """K7 (one kernel for WeatherModel.load after load_weather) against the oracle restatement, specific and relative humidity,
    with and without the extra level at zmin; then the processed cube drives the zenith path."""
    from oracle import weather as ow
    from raider_b200.weather_prep import process_weather
    zs, p, t, q = _native_columns()
    zlevels = np.concatenate([[-200.0, -50.0, 0.0, 20.0], 50.0 + 44000.0 * (np.arange(1, 60) / 59.0) ** 1.7])
    for hum, kind, zl, zmin in ((q, 'q', zlevels, -100.0), (60.0 + 30.0 * np.sin(zs / 3000.0), 'rh', zlevels[2:], -100.0)):
        got = process_weather(zs, p, t, hum, zl, humidity_type=kind, zmin=zmin, keep_pte=True)
        want = ow.process(zs, p, t, hum, zl, 0.776, 0.233, 3.75e3, humidity_type=kind, zmin=zmin)
        assert np.array_equal(got['z'], want['z']) and got['wet'].shape == want['wet'].shape and got['wet'].dtype == np.float32
        assert got['z'].size == zl.size + (1 if zmin < zl[0] else 0)
        for k in ('p', 't', 'e', 'wet', 'hydro', 'wet_total', 'hydro_total'):
            a, b = got[k].astype(np.float64), np.asarray(want[k], dtype=np.float64)
            assert np.array_equal(np.isnan(a), np.isnan(b)), k
            # float32 fields; exp() of the two libraries may differ in the last float32 bit of svp
            assert np.allclose(a, b, rtol=3e-6, atol=1e-30), (k, np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))
        assert np.isfinite(got['wet']).all() and got['t'].max() > 1e15  # levels above the column top: fillna3D's 1e16 in t
    # the reference's small known answer (test/test_weather_model.py:178-211) through the kernel: q = 0 -> e = 0; p and t interpolate
    zs_s = np.array([[[1., 2.], [0.9, 1.1]], [[1., 2.6], [1.1, 2.3]]])
    p_s = np.arange(8).reshape(2, 2, 2).astype(float)
    got = process_weather(zs_s, p_s, p_s * 2 + 300.0, np.zeros_like(p_s), np.array([1.0, 2.0]), zmin=5.0, keep_pte=True)
    nan = np.nan
    interp = np.array([[[0, nan], [2.5, nan]], [[4., 4.625], [nan, 6.75]]])
    filled_p = np.where(np.isnan(interp), 0.0, interp)
    filled_p[1, 1, 0] = 6.75  # leading NaN takes the first valid value
    assert np.allclose(np.moveaxis(got['p'], 0, 2), filled_p, rtol=0, atol=0)
