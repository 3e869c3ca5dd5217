"""Round-2 GPU tests: the fused step (device-side plan), the thin-layer / TMA-staged integrator, the nParts knife-edge guard, the
reference's own golden through the real HRRR cube, and the float64 totals.  Same tolerances as tests/test_gpu_parity.py.
"""
import logging

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_F64_M = 1e-6


@pytest.fixture(scope='module')
def gpu(lib):
    if lib.rdr_device_count() < 1:
        pytest.fail('no CUDA device visible to libraider_b200.so on a box that runs -m gpu tests')
    return lib


def _cfg(n, posting, **kw):
    from raider_b200 import synthetic as syn
    cfg = syn.config_c2(n=n, **kw)
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, n, n, posting)
    return cfg


def _trace(cube, cfg, enu, S=None, **kw):
    from raider_b200 import _lib
    ny, nx = cfg['ypts'].size, cfg['xpts'].size
    w, h = _lib.pinned_empty((ny, nx)), _lib.pinned_empty((ny, nx))
    info = cube.trace(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, float(cfg['zpts'][0]), cfg['zref'],
                      cfg['max_segment_length'] if S is None else S, w, h, **kw)
    return np.array(w), np.array(h), info


def _enu(inc=30.0, hd=-168.0):
    from raider_b200.losreader import inc_hd_to_enu
    return np.ascontiguousarray(inc_hd_to_enu(np.float64(inc), np.float64(hd)))


# ---------------------------------------------------------------------------------------- fused step == unfused step
@pytest.mark.parametrize('table', [None, 'ml145', 'hrrr57'])
def test_fused_step_equals_host_planned_step(gpu, table):
    """K0 -> k_plan -> K3 with the plan built on the device vs K0 -> host -> K3 with the maxima going through the host (the old
    ABI): same kernels, same plan -> bitwise the same maps, step counts and maxima; and both match the oracle."""
    from oracle import raytrace as rt
    from raider_b200.engine import DeviceCube
    cfg = _cfg(32, 0.02, **({'table': table} if table else {}))
    cube = DeviceCube.from_dict(cfg['cube'])
    w, h, info = _trace(cube, cfg, _enu())
    # the unfused route: identity hooks stand in for a one-rank reduction
    w2, h2, info2 = _trace(cube, cfg, _enu(), reduce_max=lambda a: a, reduce_sum=lambda a: a)
    assert np.array_equal(w, w2) and np.array_equal(h, h2)
    assert np.array_equal(info.nparts, info2.nparts) and np.array_equal(info.maxlen, info2.maxlen)
    assert info.clamp_low_first == info2.clamp_low_first and info.n_rays == info2.n_rays == 32 * 32
    crs, st = rt.GeographicCRS(), {}
    want = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], rt.FixedIncidenceLOS(30.0, -168.0), crs, crs, list(rt.get_interpolators(cfg['cube'])),
                             MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'], stats=st)
    assert np.array_equal(info.nparts, st['nParts'][0])                      # the oracle's OWN maxima: bit-exact step counts
    assert np.abs(info.maxlen - st['maxlen'][0]).max() < 1e-7
    assert np.abs(w - want[0][0]).max() < 1e-9 and np.abs(h - want[1][0]).max() < 1e-9
    if table in ('ml145', 'hrrr57'):
        assert info.k_split >= 16    # the thin-layer kernel took the lower part of the model


def test_thin_layer_kernel_staged_vs_unstaged_vs_quadrature(gpu, monkeypatch):
    """The thin-layer kernel with TMA-staged record columns, the same kernel reading the records from global memory, and the
    quadrature kernel alone (RDR_K3_THIN_MIN=0): the first two run the same arithmetic (bitwise equal), the third agrees to 1e-10 m."""
    from raider_b200.engine import DeviceCube
    cfg = _cfg(64, 0.01, table='ml145')
    cube = DeviceCube.from_dict(cfg['cube'])
    w, h, info = _trace(cube, cfg, _enu(37.0, 15.0))
    assert info.k_split >= 16 and info.staged_passes > 0
    monkeypatch.setenv('RDR_K3_THIN_STAGE', '0')
    w0, h0, info0 = _trace(cube, cfg, _enu(37.0, 15.0))
    assert info0.staged_passes == 0 and np.array_equal(w, w0) and np.array_equal(h, h0)
    monkeypatch.delenv('RDR_K3_THIN_STAGE')
    monkeypatch.setenv('RDR_K3_THIN_MIN', '0')
    wq, hq, infoq = _trace(cube, cfg, _enu(37.0, 15.0))
    assert infoq.k_split == 0 and np.array_equal(info.nparts, infoq.nparts)
    assert np.abs(w - wq).max() < 1e-10 and np.abs(h - hq).max() < 1e-10


@pytest.mark.parametrize('inc', [20.0, 45.0, 65.0])
def test_k0_layer_top_polynomial_on_the_references_145_node_table(gpu, monkeypatch, inc):
    """K0's three forms on the reference's own 145-node table (80 km, from the committed ERA-5 fixture): `poly` (default: h(t) as a
    septic, the layer tops as one degree-7 polynomial in z), `iter` (septic, three iterates per layer) and `exact` (the reference's
    iterates on PROJ-form heights).  Layer maxima within 1e-7 m of each other (measured 2e-8), nParts identical and equal to the
    oracle's from its own maxima, delays within the contract of each other and of the oracle."""
    from pathlib import Path
    from oracle import raytrace as rt
    from raider_b200 import synthetic as syn
    from raider_b200.engine import DeviceCube
    zs = np.load(Path(__file__).resolve().parent / 'golden' / 'era5_slant_ref.npz')['z']
    n = 24
    xp, yp = syn.raster(33.5, -117.8, n, n, 0.02)
    xs, ys = syn.cube_axes_around(xp, yp, pad_deg=2.5 if inc > 60 else 1.2)
    cube_d = syn.make_cube(ys, xs, zs, totals=False)
    cfg = {'cube': cube_d, 'xpts': xp, 'ypts': yp, 'zpts': np.array([0.0]), 'zref': float(zs[-1] - 1), 'max_segment_length': 1000.0}
    cube = DeviceCube.from_dict(cube_d)
    res = {}
    for mode in ('poly', 'iter', 'exact'):
        monkeypatch.setenv('RDR_K0_MODE', mode)
        res[mode] = _trace(cube, cfg, _enu(inc, -168.0))
    monkeypatch.delenv('RDR_K0_MODE')
    st = {}
    crs = rt.GeographicCRS()
    want = rt.build_cube_ray(xp, yp, cfg['zpts'], rt.FixedIncidenceLOS(inc, -168.0), crs, crs, list(rt.get_interpolators(cube_d)),
                             MAX_TROPO_HEIGHT=cfg['zref'], stats=st)
    assert not np.isnan(want[0]).any() and st['nParts'][0].size == 137
    # from ~58 deg on, the three iterates of losreader.py:720-733 overshoot the default zref (1 m below the 80 km top) by more than
    # that metre on EVERY pixel: the reference takes the last sample at max(z) (delay.py:310-311), and so does the device
    assert all(info.clamp_high_last == (inc > 60.0) for _, _, info in res.values())
    for mode, (w, h, info) in res.items():
        assert np.array_equal(info.nparts, st['nParts'][0]), mode
        assert np.abs(info.maxlen - res['exact'][2].maxlen).max() < 1e-7, mode
        assert np.abs(w - want[0][0]).max() < TOL_F64_M and np.abs(h - want[1][0]).max() < TOL_F64_M, mode
        assert np.abs(h - res['exact'][1]).max() < 1e-9, mode


def test_thin_layer_kernel_on_a_km_scale_grid(gpu):
    """A 3-km Lambert cube (C3 shape, 57-node table): the rays cross a horizontal cell every layer or two, the staged footprint is
    several columns wide; against the oracle."""
    from oracle import raytrace as rt
    from raider_b200 import synthetic as syn
    from raider_b200.losreader import Raytracing
    from raider_b200.delay import _build_cube_ray
    from raider_b200.delayFcns import getInterpolators
    c3 = syn.config_c3(ny=32, nx=32)
    ifs = getInterpolators(c3['cube'])
    out = _build_cube_ray(c3['xpts'], c3['ypts'], c3['zpts'], Raytracing(incidence=37.0, heading=-168.0), c3['crs'], 4326, list(ifs),
                          MAX_SEGMENT_LENGTH=c3['max_segment_length'], MAX_TROPO_HEIGHT=c3['zref'])
    info = ifs[0].cube.last_info[0]
    lcc = rt.LambertCRS(**c3['crs'].args)
    st = {}
    want = rt.build_cube_ray(c3['xpts'], c3['ypts'], c3['zpts'], rt.FixedIncidenceLOS(37.0, -168.0), lcc, rt.GeographicCRS(),
                             list(rt.get_interpolators(c3['cube'])), MAX_SEGMENT_LENGTH=c3['max_segment_length'], MAX_TROPO_HEIGHT=c3['zref'], stats=st)
    assert np.array_equal(info.nparts, st['nParts'][0])
    assert np.abs(out[0] - want[0]).max() < 1e-9 and np.abs(out[1] - want[1]).max() < 1e-9


# ---------------------------------------------------------------------------------------- the integer contract
def test_nparts_knife_edge_is_detected_and_redone(gpu, caplog):
    """MAX_SEGMENT_LENGTH chosen so that one layer's maximum is a whole number of segments to within 1e-9 m: the plan flags it,
    nothing is integrated with the default K0, the step is redone with the exact (Bowring) K0 and reported (detect,
    do not hide).  A segment length a hair away from the edge takes the normal route and gives the oracle's step counts."""
    from oracle import raytrace as rt
    from raider_b200.engine import DeviceCube
    cfg = _cfg(16, 0.05)
    cube = DeviceCube.from_dict(cfg['cube'])
    _, _, base = _trace(cube, cfg, _enu())
    k = 20
    S_edge = float(base.maxlen[k]) / 3.0          # layer k: maxlen / S = 3 to the last bits
    with caplog.at_level(logging.WARNING, logger='raider_b200'):
        w, h, info = _trace(cube, cfg, _enu(), S=S_edge)
    assert info.knife_edge_redo and any('knife edge' in r.message for r in caplog.records)
    assert np.isfinite(w).all() and int(info.nparts[k]) in (4, 5)
    crs = rt.GeographicCRS()
    for S in (S_edge * (1 + 1e-5), S_edge * (1 - 1e-5)):      # 10x the guard band away: the normal route
        st = {}
        want = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], rt.FixedIncidenceLOS(30.0, -168.0), crs, crs, list(rt.get_interpolators(cfg['cube'])),
                                 MAX_SEGMENT_LENGTH=S, MAX_TROPO_HEIGHT=cfg['zref'], stats=st)
        w, h, info = _trace(cube, cfg, _enu(), S=S)
        assert not info.knife_edge_redo and np.array_equal(info.nparts, st['nParts'][0])
        assert np.abs(w - want[0][0]).max() < 1e-9


def test_all_nan_is_decided_on_the_whole_raster(gpu):
    """delay.py:279-280: ValueError only when EVERY ray length of the raster is NaN.  A row tile (or a rank's block) whose rays are
    all NaN is not an error by itself: those pixels come out NaN."""
    from raider_b200 import _lib
    from raider_b200.engine import DeviceCube
    cfg = _cfg(8, 0.05)
    cube = DeviceCube.from_dict(cfg['cube'])
    ny, nx = 8, 8
    xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
    from oracle import geodesy
    enu = geodesy.inc_hd_to_enu(np.float64(30.0), np.float64(-168.0))
    los = geodesy.enu2ecef(enu[0], enu[1], enu[2], yy, xx, 0 * yy).reshape(-1, 3)
    los[: 4 * nx] = np.nan                                   # the upper half of the raster has no look vectors
    w, h = np.empty((ny, nx)), np.empty((ny, nx))
    info = cube.trace(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ARRAY, np.ascontiguousarray(los), 0.0, cfg['zref'], 225.0, w, h)
    assert info.n_nan_rays == 4 * nx and np.isnan(w[:4]).all() and np.isfinite(w[4:]).all()
    # the same raster walked in two row tiles: the first tile is all NaN, the run completes with the same result
    w2, h2 = np.empty((ny, nx)), np.empty((ny, nx))
    cube.trace(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ARRAY, np.ascontiguousarray(los), 0.0, cfg['zref'], 225.0, w2, h2,
               max_t_bytes=8 * cube.grid[2].size * nx * 4)
    assert np.array_equal(w, w2, equal_nan=True) and np.array_equal(h, h2, equal_nan=True)
    los[:] = np.nan
    with pytest.raises(ValueError, match='geo2rdr did not converge'):
        cube.trace(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ARRAY, np.ascontiguousarray(los), 0.0, cfg['zref'], 225.0, w, h)


def test_many_levels_need_the_large_shared_memory_opt_in(gpu):
    """A model with 700 levels: the layer records + z table exceed the 48 KB default of dynamic shared memory."""
    from oracle import raytrace as rt
    from raider_b200 import synthetic as syn
    from raider_b200.engine import DeviceCube
    xp, yp = syn.raster(34.0, -118.0, 8, 8, 0.05)
    xs, ys = syn.cube_axes_around(xp, yp)
    zs = np.linspace(-500.0, 40000.0, 700)
    cube_d = syn.make_cube(ys, xs, zs, totals=False)
    cfg = {'cube': cube_d, 'xpts': xp, 'ypts': yp, 'zpts': np.array([0.0]), 'zref': float(zs[-1] - 1), 'max_segment_length': 1000.0}
    w, h, info = _trace(DeviceCube.from_dict(cube_d), cfg, _enu())
    crs = rt.GeographicCRS()
    want = rt.build_cube_ray(xp, yp, cfg['zpts'], rt.FixedIncidenceLOS(30.0, -168.0), crs, crs, list(rt.get_interpolators(cube_d)), MAX_TROPO_HEIGHT=cfg['zref'])
    assert info.n_layers > 600 and np.abs(w - want[0][0]).max() < 1e-9 and np.abs(h - want[1][0]).max() < 1e-9


# ---------------------------------------------------------------------------------------- the reference's own golden, real data
def test_reference_golden_hrrr_ztd_on_device(gpu):
    """test/test_HRRR_ztd.py:18: hydro 2.2622863 m / wet 0.0361021 m at (36.84 N, 91.84 W, 0 m) from the reference's processed HRRR
    cube (a crop of it travels as tests/golden/hrrr_ztd_ref.npz, written from the HDF5 file by raider_b200.hdf5_lite) -- through
    _build_cube on the device: float64 totals (hi + lo float32 parts), Lambert model CRS, geographic query points."""
    from pathlib import Path
    from raider_b200.crs import LambertConformalSphere
    from raider_b200.delay import _build_cube
    from raider_b200.delayFcns import getInterpolators
    fx = np.load(Path(__file__).resolve().parent / 'golden' / 'hrrr_ztd_ref.npz')
    lcc = LambertConformalSphere(*fx['lcc'])
    cube = {k: fx[k] for k in ('x', 'y', 'z', 'wet', 'hydro', 'wet_total', 'hydro_total')}
    cube['crs'] = lcc
    ifs = getInterpolators(cube, 'total')
    px, py, pz = fx['gold_point']
    out = _build_cube(np.array([px]), np.array([py]), np.array([pz]), lcc, 4326, list(ifs))
    np.testing.assert_almost_equal(fx['gold_hydro_wet'], [out[1][0, 0, 0], out[0][0, 0, 0]])     # 7 decimals, as the reference asserts
    # ... and the device path is scipy's to double rounding on the float64 totals (hi + lo staging)
    from oracle import raytrace as rt
    xs = np.linspace(px - 0.05, px + 0.05, 9)
    ys = np.linspace(py - 0.05, py + 0.05, 7)
    zs = np.array([0.0, 50.0, 100.0, 500.0, 1000.0])
    want = rt.build_cube(xs, ys, zs, rt.LambertCRS(**dict(zip(('lat_1', 'lat_2', 'lat_0', 'lon_0', 'R', 'x_0', 'y_0'), (float(v) for v in fx['lcc'])))), rt.GeographicCRS(), list(rt.get_interpolators(cube, 'total')))
    got = _build_cube(xs, ys, zs, lcc, 4326, list(getInterpolators(cube, 'total')))
    assert np.abs(got[0] - want[0]).max() < 1e-12 and np.abs(got[1] - want[1]).max() < 1e-12


def test_reference_goldens_of_test_slant_on_device(gpu):
    """test/test_slant.py:49 (2.333865144 m) and :99 (2.97711681 m) through the CUDA path, on the reference's own ERA-5 cube (145
    z-nodes, real refractivity), the AOI grid the reference builds (90 x 90 x 4 heights) and a Sentinel-1 precise orbit: LOS solved
    on the device (geo2rdr), K0, K3.  The whole cubes are compared with what the REFERENCE'S OWN PYTHON produced in the build
    container (tests/golden/make_golden_era5_slant.py; fixture tests/golden/era5_slant_ref.npz), tolerance 1e-6 m."""
    import datetime as dt
    from pathlib import Path
    from raider_b200.delay import _build_cube, _build_cube_ray
    from raider_b200.delayFcns import getInterpolators
    from raider_b200.losreader import Raytracing
    gold = Path(__file__).resolve().parent / 'golden'
    fx = np.load(gold / 'era5_slant_ref.npz')
    cube = {k: fx[k] for k in ('x', 'y', 'z', 'wet', 'hydro', 'wet_total', 'hydro_total')}
    iy, ix = fx['gold_index']
    zw, zh = _build_cube(fx['xpts'], fx['ypts'], fx['zpts'], 4326, 4326, list(getInterpolators(cube, 'total')))
    np.testing.assert_almost_equal(float(fx['gold_std']), (zw + zh)[0, iy, ix])           # 7 decimals, as the reference asserts
    assert np.abs(zw - fx['ref_ztd_wet']).max() < 1e-12 and np.abs(zh - fx['ref_ztd_hydro']).max() < 1e-12
    los = Raytracing(filename=str(gold / 'orbit_S1B_20200130_sv.txt'), time=dt.datetime.fromisoformat(str(fx['time'])))
    assert los.getSensorDirection() == 'desc'
    ifs = getInterpolators(cube)
    rw, rh = _build_cube_ray(fx['xpts'], fx['ypts'], fx['zpts'], los, 4326, 4326, list(ifs), MAX_TROPO_HEIGHT=float(fx['zref']))
    np.testing.assert_almost_equal(float(fx['gold_ray']), (rw + rh)[0, iy, ix])
    assert not np.isnan(rw).any() and not np.isnan(rh).any()
    assert np.abs(rw - fx['ref_ray_wet']).max() < TOL_F64_M and np.abs(rh - fx['ref_ray_hydro']).max() < TOL_F64_M
    info = ifs[0].cube.last_info
    assert len(info) == 4 and all(i.n_layers > 100 for i in info)      # the 145-node production table


def test_ray_tracing_through_the_references_hrrr_cube(gpu):
    """Slant delays through the REAL refractivity fields of the reference's HRRR cube (3 km Lambert grid, the 57-node table of
    models/model_levels.py:517 as the file holds it) against the oracle: real data, real level table, projected model CRS."""
    from pathlib import Path
    from oracle import raytrace as rt
    from raider_b200.crs import LambertConformalSphere
    from raider_b200.delay import _build_cube_ray
    from raider_b200.delayFcns import getInterpolators
    from raider_b200.losreader import Raytracing
    fx = np.load(Path(__file__).resolve().parent / 'golden' / 'hrrr_ztd_ref.npz')
    lcc = LambertConformalSphere(*fx['lcc'])
    cube = {k: fx[k] for k in ('x', 'y', 'z', 'wet', 'hydro')}
    cube['crs'] = lcc
    px, py, _ = fx['gold_point']
    xs = np.linspace(px - 0.03, px + 0.03, 16)
    ys = np.linspace(py - 0.03, py + 0.03, 12)
    zpts = np.array([0.0, 500.0])
    zref = float(cube['z'][-1] - 1)
    ifs = getInterpolators(cube)
    out = _build_cube_ray(xs, ys, zpts, Raytracing(incidence=25.0, heading=-168.0), lcc, 4326, list(ifs), MAX_TROPO_HEIGHT=zref)
    st = {}
    want = rt.build_cube_ray(xs, ys, zpts, rt.FixedIncidenceLOS(25.0, -168.0), rt.LambertCRS(**dict(zip(('lat_1', 'lat_2', 'lat_0', 'lon_0', 'R', 'x_0', 'y_0'), (float(v) for v in fx['lcc'])))), rt.GeographicCRS(),
                             list(rt.get_interpolators(cube)), MAX_TROPO_HEIGHT=zref, stats=st)
    info = ifs[0].cube.last_info
    assert np.array_equal(info[0].nparts, st['nParts'][0]) and np.array_equal(info[1].nparts, st['nParts'][1])
    assert np.isfinite(want[0]).all()
    assert np.abs(out[0] - want[0]).max() < TOL_F64_M and np.abs(out[1] - want[1]).max() < TOL_F64_M
    assert np.abs(out[1] - want[1]).max() < 1e-9


def test_caller_supplied_output_arrays_accumulate(gpu):
    """outputArrs given (delay.py:245-248,323): in-place +=, through page-locked scratch the kernel writes directly."""
    from raider_b200.delay import _build_cube_ray
    from raider_b200.delayFcns import getInterpolators
    from raider_b200.losreader import Raytracing
    cfg = _cfg(12, 0.05)
    ifs = getInterpolators(cfg['cube'])
    los = Raytracing(incidence=30.0, heading=-168.0)
    fresh = _build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], los, 4326, 4326, list(ifs), MAX_SEGMENT_LENGTH=225.0, MAX_TROPO_HEIGHT=cfg['zref'])
    seed = [np.full((1, 12, 12), 1.5), np.full((1, 12, 12), -2.0)]
    mine = [a.copy() for a in seed]
    assert _build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], los, 4326, 4326, list(ifs), outputArrs=mine, MAX_SEGMENT_LENGTH=225.0,
                           MAX_TROPO_HEIGHT=cfg['zref']) is None
    assert np.array_equal(mine[0], seed[0] + fresh[0]) and np.array_equal(mine[1], seed[1] + fresh[1])


def test_upper_clamp_of_the_last_sample(gpu):
    """delay.py:310-311: when the top of the top layer lies above max(z) on EVERY pixel (steep rays, zref at its default of 1 m below
    the model top) the reference samples it at max(z).  K0 counts it, the plan decides it globally, the PROJ-form kernel applies it
    (the polynomial kernels hand such rays over); the hosted two-call form and the fused step agree; a raster where only SOME pixels
    overshoot gets NaN on those, like the reference (the `.all()` is false)."""
    from pathlib import Path
    from oracle import raytrace as rt
    from raider_b200 import _lib, synthetic as syn
    from raider_b200.engine import DeviceCube
    zs = np.load(Path(__file__).resolve().parent / 'golden' / 'era5_slant_ref.npz')['z']
    n = 16
    xp, yp = syn.raster(33.5, -117.8, n, n, 0.02)
    xs, ys = syn.cube_axes_around(xp, yp, pad_deg=3.0)
    cube_d = syn.make_cube(ys, xs, zs, totals=False)
    cfg = {'cube': cube_d, 'xpts': xp, 'ypts': yp, 'zpts': np.array([0.0]), 'zref': float(zs[-1] - 1), 'max_segment_length': 1000.0}
    cube = DeviceCube.from_dict(cube_d)
    crs = rt.GeographicCRS()
    ifs = list(rt.get_interpolators(cube_d))
    # (a) every pixel overshoots
    want = rt.build_cube_ray(xp, yp, cfg['zpts'], rt.FixedIncidenceLOS(66.0, -168.0), crs, crs, ifs, MAX_TROPO_HEIGHT=cfg['zref'])
    assert not np.isnan(want[0]).any()
    w, h, info = _trace(cube, cfg, _enu(66.0, -168.0))
    assert info.clamp_high_last and not info.clamp_low_first and info.oob_above == 0
    assert np.abs(w - want[0][0]).max() < TOL_F64_M and np.abs(h - want[1][0]).max() < TOL_F64_M
    assert np.abs(h - want[1][0]).max() < 1e-9
    # the two-call (hosted) form decides the same from counts_out[4] (identity hooks stand in for a one-rank reduction), also in row tiles
    w2, h2, info2 = _trace(cube, cfg, _enu(66.0, -168.0), reduce_max=lambda a: a, reduce_sum=lambda a: a)
    assert info2.clamp_high_last and np.array_equal(w2, w) and np.array_equal(h2, h)
    w3, h3, info3 = _trace(cube, cfg, _enu(66.0, -168.0), max_t_bytes=8 * zs.size * n * 5)
    assert info3.tiles > 1 and info3.clamp_high_last and np.array_equal(w3, w) and np.array_equal(h3, h)
    # (b) an explicit LOS array: half the raster at 66 deg, half at 30 deg -> the predicate is false, the steep half is NaN (reference too)
    from oracle import geodesy
    inc = np.where(np.arange(n)[None, :] < n // 2, 66.0, 30.0) * np.ones((n, 1))
    xx, yy = np.meshgrid(xp, yp)
    enu = geodesy.inc_hd_to_enu(inc, np.full_like(inc, -168.0))
    look = geodesy.enu2ecef(enu[..., 0], enu[..., 1], enu[..., 2], yy, xx, np.zeros_like(yy))
    want = rt.build_cube_ray(xp, yp, cfg['zpts'], rt.ArrayLOS(look), crs, crs, ifs, MAX_TROPO_HEIGHT=cfg['zref'])
    ww, hh = _lib.pinned_empty((n, n)), _lib.pinned_empty((n, n))
    info = cube.trace(_lib.GEOM_GRID, xp, yp, n, n, _lib.LOS_ARRAY, np.ascontiguousarray(look.reshape(-1, 3)), 0.0, cfg['zref'], 1000.0, ww, hh)
    assert not info.clamp_high_last
    assert np.isnan(want[0][0][:, : n // 2]).all() and not np.isnan(want[0][0][:, n // 2:]).any()
    assert np.array_equal(np.isnan(np.array(ww)), np.isnan(want[0][0]))
    assert np.nanmax(np.abs(np.array(hh) - want[1][0])) < TOL_F64_M


def test_reference_golden_of_test_gnss_intersect_on_device(gpu):
    """test/test_intersect.py:104 (TORP 2.34514 m total zenith delay, 4 decimals) through the product's tropo_delay in station (point)
    mode: ZTD cube on the reference's AOI grid at the model's 145 z levels (device, float64 totals as hi + lo), then the cube
    re-interpolated at the stations (device).  Against the values the reference's own Python produced (tests/golden/era5_gnss_ref.npz)."""
    import datetime as dt
    from pathlib import Path
    from raider_b200.delay import tropo_delay
    from raider_b200.llreader import Points
    from raider_b200.losreader import Zenith
    gold = Path(__file__).resolve().parent / 'golden'
    fx, cx = np.load(gold / 'era5_gnss_ref.npz'), np.load(gold / 'era5_slant_ref.npz')
    cube = {k: cx[k] for k in ('x', 'y', 'z', 'wet', 'hydro', 'wet_total', 'hydro_total')}
    aoi = Points(fx['lats'], fx['lons'], fx['hgts'])
    aoi.xpts, aoi.ypts = fx['xpts'], fx['ypts']          # the grid the reference's StationFile AOI builds (cli/raider.py:257-260)
    wet, hydro = tropo_delay(dt.datetime(2020, 1, 30, 13, 52, 45), cube, aoi, Zenith())
    np.testing.assert_almost_equal((wet + hydro)[int(fx['gold_index'])], float(fx['gold_total']), decimal=4)
    assert np.abs(wet - fx['ref_wet']).max() < 1e-11 and np.abs(hydro - fx['ref_hydro']).max() < 1e-11
