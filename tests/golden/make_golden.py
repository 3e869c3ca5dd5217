"""Generate the committed golden vectors under tests/golden/ -- run HERE (CPU container), never on the GPU box.

    python tests/golden/make_golden.py

Sources of truth, in order of authority:
  1. the reference's own native code compiled from /root/reference into oracle/_ref (oracle/build_ref.py):
     ``interpolate``, ``interpolate_along_axis``, ``makePoints{0..3}D``;
  2. the reference's golden file test/test_result_makePoints3D.txt (checked bit-for-bit against (1) here);
  3. the installed scipy RegularGridInterpolator (the third-party sampler the delay path really calls);
  4. for the ray tracer: the reference's OWN Python (RAiDER.delay._build_cube_ray, RAiDER.losreader.build_ray /
     getTopOfAtmosphere, RAiDER.delayFcns.getInterpolators) imported unmodified from /root/reference by oracle/refpy.py,
     with stand-ins only for the packages that are not installable offline (pyproj -> oracle.geodesy's restatement of
     PROJ's cart / lcc, xarray -> an in-memory Dataset).  raytrace.npz and the ray part of geodesy.npz are written from
     the REFERENCE functions' outputs; the NumPy restatement in oracle/raytrace.py is asserted bitwise equal on the way.

Every vector is seeded; the GPU tests compare against these files, so they hold even where /root/reference and the
oracle's dependencies are absent.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent
REFERENCE = Path('/root/reference')

from oracle import build_ref, geodesy, interp as ointerp, raytrace as rt, refpy  # noqa: E402
from raider_b200 import synthetic as syn  # noqa: E402

REF = refpy.load()


def golden_makepoints(ref_mp):
    sp = np.zeros((3, 3, 3, 3))
    sp[:, :, 1, 2] = 10
    sp[:, :, 2, 2] = 100
    slv = np.zeros((3, 3, 3, 3))
    slv[0, :, :, 2] = 1
    slv[1, :, :, 1] = 1
    slv[2, :, :, 0] = 1
    res = ref_mp.makePoints3D(100.0, sp, slv, 5.0)
    txt = np.loadtxt(REFERENCE / 'test' / 'test_result_makePoints3D.txt').reshape((3, 3, 3, 3, 20))
    assert np.array_equal(res, txt), 'compiled reference makePoints3D != reference golden file'
    rng = np.random.default_rng(11)
    sp_r = rng.normal(scale=6.4e6, size=(5, 7, 3))
    slv_r = rng.normal(size=(5, 7, 3))
    slv_r /= np.linalg.norm(slv_r, axis=-1, keepdims=True)
    res_r = ref_mp.makePoints2D(5000.0, sp_r, slv_r, 15.0)
    counts = np.array([[L, s, ref_mp.makePoints0D(L, np.zeros(3), np.ones(3), s).shape[1]]
                       for L, s in [(100, 5), (100, 15), (101, 5), (1000, 5), (20, 5), (0.3, 0.1), (1.0, 0.1), (7.5, 2.5), (3.0, 7.0)]], dtype=np.float64)
    np.savez_compressed(OUT / 'makepoints.npz', sp3=sp, slv3=slv, out3=res, sp2=sp_r, slv2=slv_r, out2=res_r, counts=counts)


def golden_interpolate(ref_it):
    rng = np.random.default_rng(1234)
    out = {}
    for nd, shape in [(1, (17,)), (2, (9, 13)), (3, (7, 11, 13)), (4, (4, 5, 6, 7))]:
        grids = [np.sort(rng.uniform(-5, 20, size=s)) for s in shape]
        vals = rng.normal(size=shape)
        n = 400
        pts = np.stack([rng.uniform(g[0] - 2.0, g[-1] + 2.0, size=n) for g in grids], axis=-1)
        # exact nodes and edges
        for d, g in enumerate(grids):
            pts[d, d] = g[0]
            pts[10 + d, d] = g[-1]
            pts[20 + d, d] = g[len(g) // 2]
        pts[30, 0] = np.nan
        out[f'nd{nd}_ngrid'] = np.array(len(grids))
        for d, g in enumerate(grids):
            out[f'nd{nd}_g{d}'] = g
        out[f'nd{nd}_vals'] = vals
        out[f'nd{nd}_pts'] = pts
        out[f'nd{nd}_fill'] = ref_it.interpolate(grids, vals, pts, fill_value=np.nan, max_threads=1)
        out[f'nd{nd}_clamp'] = ref_it.interpolate(grids, vals, pts, max_threads=1)
        # the oracle restatement must already agree bit-for-bit
        assert np.array_equal(ointerp.interpolate(grids, vals, pts, fill_value=np.nan), out[f'nd{nd}_fill'], equal_nan=True)
        assert np.array_equal(ointerp.interpolate(grids, vals, pts), out[f'nd{nd}_clamp'], equal_nan=True)
    # along axis: per-column grids, axis 2 and axis 1 (the _uniform_in_z shape, weatherModel.py:617-619)
    zs = np.sort(rng.uniform(0, 40000, size=(6, 5, 30)), axis=2)
    f = rng.normal(size=zs.shape)
    new = np.sort(rng.uniform(-500, 41000, size=(6, 5, 41)), axis=2)
    new[..., 3] = zs[..., 4]
    out['ax_x'], out['ax_y'], out['ax_new'] = zs, f, new
    out['ax_fill'] = ref_it.interpolate_along_axis(zs, f, new, axis=2, fill_value=np.nan, max_threads=2)
    out['ax_clamp'] = ref_it.interpolate_along_axis(zs, f, new, axis=2, max_threads=1)
    x1 = np.ascontiguousarray(np.moveaxis(zs, 2, 1))
    y1 = np.ascontiguousarray(np.moveaxis(f, 2, 1))
    n1 = np.ascontiguousarray(np.moveaxis(new, 2, 1))
    out['ax1_fill'] = ref_it.interpolate_along_axis(x1, y1, n1, axis=1, fill_value=np.nan)
    assert np.array_equal(ointerp.interpolate_along_axis(zs, f, new, axis=2, fill_value=np.nan), out['ax_fill'], equal_nan=True)
    np.savez_compressed(OUT / 'interpolate.npz', **out)


def golden_scipy():
    """scipy RGI exactly as delayFcns.py:55-56 configures it, incl. the boundary table of SURVEY.md Appendix B."""
    from scipy.interpolate import RegularGridInterpolator as RGI
    rng = np.random.default_rng(77)
    ys, xs, zs = np.linspace(30, 36, 9), np.linspace(-120, -114, 11), syn.z_levels(13)
    wet = rng.normal(size=(13, 9, 11)).astype(np.float32)
    hydro = rng.normal(size=(13, 9, 11)).astype(np.float32)
    pts = np.stack([rng.uniform(29.9, 36.1, 3000), rng.uniform(-120.1, -113.9, 3000), rng.uniform(-600, 48100, 3000)], axis=-1)
    pts[0] = [30, -120, zs[0]]
    pts[1] = [36, -114, zs[-1]]
    pts[2] = [36, -117.3, 100.0]
    pts[3] = [33, -114, zs[5]]
    pts[4] = [30 - 1e-9, -117, 0]
    pts[5] = [33, -117, zs[-1] + 1e-7]
    pts[6] = [np.nan, -117, 0]
    pts[7] = [ys[3], xs[4], zs[6]]
    w = RGI((ys, xs, zs), wet.transpose(1, 2, 0), fill_value=np.nan, bounds_error=False)(pts)
    h = RGI((ys, xs, zs), hydro.transpose(1, 2, 0), fill_value=np.nan, bounds_error=False)(pts)
    np.savez_compressed(OUT / 'scipy_sample.npz', ys=ys, xs=xs, zs=zs, wet=wet, hydro=hydro, pts=pts, out_wet=w, out_hydro=h)


def _ref_crs(model_crs):
    """oracle CRS object -> the (stand-in) pyproj CRS the reference functions take."""
    if model_crs is None or model_crs.kind == 0:
        return REF.CRS.from_epsg(4326)
    n, c, rho0, lam0, R, x0, y0 = model_crs.params()
    l = model_crs.lcc
    return REF.CRS(dict(proj='lcc', lat_1=l.lat_1, lat_2=l.lat_2, lat_0=l.lat_0, lon_0=l.lon_0, a=R, b=R, x_0=x0, y_0=y0))


def _ref_los(los):
    if isinstance(los, rt.ZenithLOS):
        return refpy.zenith_los(REF)
    if isinstance(los, rt.FixedIncidenceLOS):
        return refpy.fixed_incidence_los(REF, los.incidence_deg, los.heading_deg)
    return refpy.array_los(los.vecs)


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def _trace(cfg, los, model_crs=None, pts_crs=None, kind='pointwise'):
    """Run the REFERENCE's _build_cube_ray (the golden) and the oracle restatement; they must agree bit for bit."""
    mcrs, pcrs = _ref_crs(model_crs), _ref_crs(pts_crs)
    zpts = np.asarray(cfg['zpts'], dtype=np.float64)
    ifs_ref = REF.delayFcns.getInterpolators(refpy.dataset(cfg['cube']), kind)
    out_ref = REF.delay._build_cube_ray(cfg['xpts'], cfg['ypts'], zpts, _ref_los(los), mcrs, pcrs, list(ifs_ref),
                                        MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
    # nParts / per-layer maxima from the reference's build_ray (delay.py:262-283 restated around the reference calls)
    st_ref = {'nParts': [], 'maxlen': []}
    xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
    for ht in zpts:
        llh = [xx, yy, np.full(yy.shape, ht)]
        if pcrs != REF.CRS.from_epsg(4326):
            llh = list(REF.Transformer.from_crs(pcrs, 4326, always_xy=True).transform(*llh))
        xyz = np.stack(REF.utilFcns.lla2ecef(llh[1], llh[0], llh[2]), axis=-1)
        lens, _, _ = REF.losreader.build_ray(ifs_ref[0].grid[2], ht, xyz, _ref_los(los).getLookVectors(ht, llh, xyz, yy), cfg['zref'])
        if lens is None:
            continue
        st_ref['maxlen'].append(lens.max((1, 2)))
        st_ref['nParts'].append(np.ceil(lens.max((1, 2)) / cfg['max_segment_length']).astype(int) + 1)

    ifs = rt.get_interpolators(cfg['cube'], kind)
    st = {}
    out = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], zpts, los, model_crs or rt.GeographicCRS(), pts_crs or rt.GeographicCRS(),
                            list(ifs), MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'], stats=st)
    assert _same(out[0], out_ref[0]) and _same(out[1], out_ref[1]), 'oracle.raytrace.build_cube_ray != reference _build_cube_ray'
    assert len(st['nParts']) == len(st_ref['nParts'])
    for a, b in zip(st['nParts'] + st['maxlen'], st_ref['nParts'] + st_ref['maxlen']):
        assert _same(np.asarray(a), np.asarray(b)), 'oracle nParts / maxima != reference'
    return out_ref, st_ref


def golden_raytrace():
    out = {}
    # (a) C2 shape, small raster: 30 deg incidence, NZ = 37, 225 m segments
    cfg = syn.config_c2(n=24)
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 24, 24, 0.08)  # spread over the full 2-degree footprint
    res, st = _trace(cfg, rt.FixedIncidenceLOS(cfg['incidence'], cfg['heading']))
    out.update(a_xpts=cfg['xpts'], a_ypts=cfg['ypts'], a_wet=res[0], a_hydro=res[1], a_nparts=st['nParts'][0], a_maxlen=st['maxlen'][0])
    # (b) 145-node table, default 1000 m segments, two output heights, 45 deg
    cfg = syn.config_c2(n=12, table='ml145')
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 12, 12, 0.15)
    cfg['zpts'] = np.array([0.0, 1500.0])
    res, st = _trace(cfg, rt.FixedIncidenceLOS(45.0, 12.0))
    out.update(b_xpts=cfg['xpts'], b_ypts=cfg['ypts'], b_wet=res[0], b_hydro=res[1], b_nparts=st['nParts'][0],
               b_nparts1=st['nParts'][1], b_maxlen=st['maxlen'][0])
    # (c) zenith ray tracing through the same cube as (a)
    cfg = syn.config_c2(n=10)
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 10, 10, 0.2)
    res, st = _trace(cfg, rt.ZenithLOS())
    out.update(c_xpts=cfg['xpts'], c_ypts=cfg['ypts'], c_wet=res[0], c_hydro=res[1], c_nparts=st['nParts'][0])
    # (d) explicit per-pixel LOS array (what Raytracing.getLookVectors returns): incidence varying 20..46 deg across the raster
    cfg = syn.config_c2(n=16)
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 16, 16, 0.1)
    xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
    inc = 20.0 + 26.0 * (xx - xx.min()) / (xx.max() - xx.min())
    enu = geodesy.inc_hd_to_enu(inc, np.full(inc.shape, -168.0))
    vecs = geodesy.enu2ecef(enu[..., 0], enu[..., 1], enu[..., 2], yy, xx, 0 * yy)
    res, st = _trace(cfg, rt.ArrayLOS(vecs))
    out.update(d_xpts=cfg['xpts'], d_ypts=cfg['ypts'], d_los=vecs, d_wet=res[0], d_hydro=res[1], d_nparts=st['nParts'][0])
    # (e) zenith/projected path (_build_cube) on the C1 cube
    c1 = syn.config_c1()
    ifs = rt.get_interpolators(c1['cube'], 'total')
    g = rt.GeographicCRS()
    zt_port = rt.build_cube(c1['xpts'][::5], c1['ypts'][::5], c1['zpts'], g, g, list(ifs))
    crs4326 = REF.CRS.from_epsg(4326)
    zt = REF.delay._build_cube(c1['xpts'][::5], c1['ypts'][::5], np.asarray(c1['zpts']), crs4326, crs4326,
                               list(REF.delayFcns.getInterpolators(refpy.dataset(c1['cube']), 'total')))
    assert _same(zt[0], zt_port[0]) and _same(zt[1], zt_port[1]), 'oracle build_cube != reference _build_cube'
    out.update(e_xpts=c1['xpts'][::5], e_ypts=c1['ypts'][::5], e_zpts=c1['zpts'], e_wet=zt[0], e_hydro=zt[1])
    # (f) HRRR-like LCC cube (3 km grid, 57-node table), geographic query raster
    lcc = rt.LambertCRS()
    cx, cy = lcc.lcc.forward(-98.0, 36.0)
    xs = cx + 3000.0 * (np.arange(60) - 30)
    ys = cy + 3000.0 * (np.arange(50) - 25)
    X, Y = np.meshgrid(xs, ys)
    lon_n, lat_n = lcc.lcc.inverse(X, Y)
    cube = syn.make_cube(ys, xs, syn.z_levels_table('hrrr57'), lat_of=lat_n, lon_of=lon_n, seed=5)
    xpts, ypts = syn.raster(36.0, -98.0, 12, 12, 0.03)
    cfgf = dict(cube=cube, xpts=xpts, ypts=ypts, zpts=np.array([200.0]), zref=float(cube['z'][-1] - 1), max_segment_length=1000.0)
    res, st = _trace(cfgf, rt.FixedIncidenceLOS(35.0, -12.0), model_crs=lcc)
    out.update(f_xs=xs, f_ys=ys, f_wet_cube=cube['wet'], f_hydro_cube=cube['hydro'], f_xpts=xpts, f_ypts=ypts, f_wet=res[0],
               f_hydro=res[1], f_nparts=st['nParts'][0], f_lcc=lcc.params())
    np.savez_compressed(OUT / 'raytrace.npz', **out)


def golden_geodesy():
    rng = np.random.default_rng(3)
    lat = np.concatenate([rng.uniform(-89.999, 89.999, 500), [0, 0, 90, -90, 45.0]])
    lon = np.concatenate([rng.uniform(-180, 180, 500), [0, 90, 0, 0, 180.0]])
    h = np.concatenate([rng.uniform(-1000, 90000, 500), [0, 0, 0, 0, 1000.0]])
    x, y, z = geodesy.lla2ecef(lat, lon, h)
    lo, la, hh = geodesy.ecef2lla(x, y, z)
    # getTopOfAtmosphere / build_ray on a handful of rays
    g = np.stack([x[:64], y[:64], z[:64]], axis=-1)
    enu = geodesy.inc_hd_to_enu(rng.uniform(0, 60, 64), rng.uniform(0, 360, 64))
    look = geodesy.enu2ecef(enu[:, 0], enu[:, 1], enu[:, 2], lat[:64], lon[:64], h[:64])
    g0 = np.stack(geodesy.lla2ecef(lat[:64], lon[:64], np.zeros(64)), axis=-1)
    zs = syn.z_levels(37)
    # the reference's own losreader.getTopOfAtmosphere / build_ray (losreader.py:706-733,772-835) write the golden
    toa10 = REF.losreader.getTopOfAtmosphere(g0, look, 30000.0)
    toa3 = REF.losreader.getTopOfAtmosphere(g0, look, 30000.0, factor=enu[:, 2])
    lens, lows, highs = REF.losreader.build_ray(zs, 0.0, g0, look, zs[-1] - 1)
    assert _same(toa10, rt.getTopOfAtmosphere(g0, look, 30000.0))
    assert _same(toa3, rt.getTopOfAtmosphere(g0, look, 30000.0, factor=enu[:, 2]))
    for a, b in zip((lens, lows, highs), rt.build_ray(zs, 0.0, g0, look, zs[-1] - 1)):
        assert _same(a, b), 'oracle build_ray != reference build_ray'
    assert _same(np.stack(REF.utilFcns.lla2ecef(lat, lon, h), -1), np.stack([x, y, z], -1))
    assert _same(REF.losreader.inc_hd_to_enu(30.0, -168.0), geodesy.inc_hd_to_enu(30.0, -168.0))
    np.savez_compressed(OUT / 'geodesy.npz', lat=lat, lon=lon, h=h, x=x, y=y, z=z, lon_back=lo, lat_back=la, h_back=hh, g0=g0, look=look,
                        cosf=enu[:, 2], toa10=toa10, toa3=toa3, zs=zs, lens=lens, lows=lows, highs=highs)


if __name__ == '__main__':
    assert build_ref.build(), 'oracle/_ref could not be built (is /root/reference present?)'
    golden_makepoints(build_ref.load('makePoints'))
    golden_interpolate(build_ref.load('interpolate'))
    golden_scipy()
    golden_raytrace()
    golden_geodesy()
    for p in sorted(OUT.glob('*.npz')):
        print(f'{p.name}: {p.stat().st_size / 1024:.1f} KiB')
