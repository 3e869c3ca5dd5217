"""Cube ingest: the dependency-free HDF5 / NetCDF-4 reader against the reference's own processed weather-model files and golden
(CPU; the reference files exist only in the build container -- the GPU box has the compact fixture tests/golden/hrrr_ztd_ref.npz).

The strongest pin here is Tier B of SURVEY 8(c): test/test_HRRR_ztd.py:18 -- the reference's zenith delays at (36.84 N, 91.84 W,
0 m) from ``HRRR_2020_01_01_T12_00_00_35N_38N_93W_90W.nc``, hydro 2.2622863 m / wet 0.0361021 m, produced by the reference with
the real PROJ / xarray / scipy stack.  Reading that HDF5 file with raider_b200.hdf5_lite and running the oracle's ``build_cube``
(our restatement of PROJ's spherical Lambert forward included) reproduces both numbers to the 7 decimals the reference asserts.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import raytrace as rt
from raider_b200 import cube_io, hdf5_lite
from raider_b200.crs import Geographic, LambertConformalSphere

REF = Path('/root/reference/test')
needs_ref = pytest.mark.skipif(not REF.exists(), reason='/root/reference is only present in the build container')
HRRR = REF / 'scenario_1' / 'HRRR_ztd_test' / 'HRRR_2020_01_01_T12_00_00_35N_38N_93W_90W.nc'
GOLD = Path(__file__).resolve().parent / 'golden'


@needs_ref
@pytest.mark.parametrize('rel,shape,crs_type', [
    ('weather_files/ERA-5_2020_01_30_T13_52_45_32N_35N_120W_115W.nc', (145, 12, 17), Geographic),
    ('weather_files/ERA-5_2019_11_17_T20_51_58_5S_2S_41W_37W.nc', (145, 10, 15), Geographic),
    ('weather_files/ERA-5_2022_08_29_T17_00_01_69N_73N_159W_152W.nc', (145, 13, 25), Geographic),
    ('gunw_test_data/weather_files/GMAO_2020_01_24_T12_00_00_32N_36N_121W_114W.nc', (145, 17, 20), Geographic),
    ('scenario_1/HRRR_ztd_test/HRRR_2020_01_01_T12_00_00_35N_38N_93W_90W.nc', (57, 50, 42), LambertConformalSphere),
    ('gunw_azimuth_test_data/weather_files/HRRR_2021_07_11_T01_00_00_33N_36N_120W_115W.nc', (57, 114, 143), LambertConformalSphere),
])
def test_reads_the_references_processed_cubes(rel, shape, crs_type):
    """Variables, shapes and dtypes as the reference writes them (weatherModel.py:659-724); the zenith totals in the file are the
    cumulative trapezoid of the refractivities in the file (weatherModel.py:389-403) -- a decoding error anywhere would break that."""
    cube = cube_io.load_cube(REF / rel)
    assert isinstance(cube['crs'], crs_type)
    for k in ('wet', 'hydro'):
        assert cube[k].shape == shape and cube[k].dtype == np.float32
    for k in ('wet_total', 'hydro_total'):
        assert cube[k].shape == shape and cube[k].dtype == np.float64
    assert cube['z'].shape == (shape[0],) and np.all(np.diff(cube['z']) > 0) and cube['z'][0] == -500.0
    assert np.all(np.diff(cube['x']) > 0) and np.all(np.diff(cube['y']) > 0)
    for f in ('wet', 'hydro'):
        want = rt.cumulative_ztd(cube[f], cube['z'])
        assert np.allclose(cube[f + '_total'], want, rtol=1e-6, atol=1e-9)
    if crs_type is LambertConformalSphere:
        assert cube['crs'].args == dict(lat_1=38.5, lat_2=38.5, lat_0=38.5, lon_0=262.5, R=6371229.0, x_0=0.0, y_0=0.0)  # models/hrrr.py:248-260


@needs_ref
def test_reference_golden_hrrr_ztd():
    """test/test_HRRR_ztd.py:18 (the reference's own golden for this cube) through hdf5_lite + the oracle's _build_cube."""
    cube = cube_io.load_cube(HRRR)
    lcc = rt.LambertCRS(**cube['crs'].args)
    out = rt.build_cube(np.array([-91.84]), np.array([36.84]), np.array([0.0]), lcc, rt.GeographicCRS(), list(rt.get_interpolators(cube, 'total')))
    np.testing.assert_almost_equal(2.2622863, out[1][0, 0, 0])   # hydro, 7 decimals as the reference asserts
    np.testing.assert_almost_equal(0.0361021, out[0][0, 0, 0])   # wet
    # the compact fixture the GPU box uses holds exactly these arrays
    fx = np.load(GOLD / 'hrrr_ztd_ref.npz')
    y0, y1, x0, x1 = fx['crop']
    assert np.array_equal(fx['x'], cube['x'][x0:x1]) and np.array_equal(fx['y'], cube['y'][y0:y1]) and np.array_equal(fx['z'], cube['z'])
    for k in ('wet', 'hydro', 'wet_total', 'hydro_total'):
        assert np.array_equal(fx[k], cube[k][:, y0:y1, x0:x1])
    # ... and the crop alone reproduces the golden as well
    crop = {k: fx[k] for k in ('x', 'y', 'z', 'wet', 'hydro', 'wet_total', 'hydro_total')}
    out = rt.build_cube(fx['gold_point'][:1], fx['gold_point'][1:2], fx['gold_point'][2:], lcc, rt.GeographicCRS(), list(rt.get_interpolators(crop, 'total')))
    np.testing.assert_almost_equal(fx['gold_hydro_wet'], [out[1][0, 0, 0], out[0][0, 0, 0]])


@needs_ref
def test_reads_groups_chunks_and_filters():
    """A GUNW product of the reference's tests: old-style groups (superblock 0), chunked + deflate + shuffle datasets, vlen strings."""
    p = REF / 'gunw_test_data' / 'S1-GUNW-D-R-071-tops-20200130_20200124-135156-34956N_32979N-PP-913f-v2_0_4.nc'
    with hdf5_lite.File(p) as f:
        g = f['science/grids/data']
        phase = g['unwrappedPhase']
        a = phase.read()
        assert a.shape == (2375, 3745) and a.dtype == np.float32 and phase.attrs['units'] == 'rad'
        lat, lon = g['latitude'].read(), g['longitude'].read()
        assert lat.shape == (2375,) and lon.shape == (3745,) and np.all(np.diff(lat) < 0) and np.all(np.diff(lon) > 0)
        assert 32.9 < lat.min() < lat.max() < 35.0
        assert np.isfinite(a).mean() > 0.5 and -40 < np.nanmin(a) < np.nanmax(a) < 40
        assert 'science/radarMetaData/inputSLC'.split('/')[1] in f['science'].keys()


@needs_ref
def test_not_hdf5_is_refused(tmp_path):
    p = tmp_path / 'x.nc'
    p.write_bytes(b'not a netcdf file at all')
    with pytest.raises(ValueError):
        cube_io.load_cube(p)
    with pytest.raises(hdf5_lite.HDF5FormatError):
        hdf5_lite.File(p)


def test_netcdf3_round_trip_keeps_the_references_dtypes(tmp_path):
    """write_cube -> load_cube: float32 refractivities, float64 totals (weatherModel.py:398-403,617-619), CRS through proj4."""
    rng = np.random.default_rng(0)
    zs = np.array([-500.0, 0.0, 700.0, 2000.0])
    cube = {'x': np.linspace(-119, -117, 5), 'y': np.linspace(33, 35, 4), 'z': zs,
            'wet': rng.normal(size=(4, 4, 5)).astype(np.float32), 'hydro': rng.normal(size=(4, 4, 5)).astype(np.float32)}
    cube['wet_total'] = rt.cumulative_ztd(cube['wet'], zs)
    cube['hydro_total'] = rt.cumulative_ztd(cube['hydro'], zs)
    back = cube_io.load_cube(cube_io.write_cube(tmp_path / 'c.nc', cube))
    for k in ('wet', 'hydro', 'wet_total', 'hydro_total', 'x', 'y', 'z'):
        assert np.array_equal(back[k], cube[k]) and back[k].dtype == cube[k].dtype
    assert isinstance(back['crs'], Geographic)
    lcc = '+proj=lcc +lat_1=38.5 +lat_2=38.5 +lat_0=38.5 +lon_0=262.5 +a=6371229 +b=6371229'
    back = cube_io.load_cube(cube_io.write_cube(tmp_path / 'l.nc', cube, proj4=lcc))
    assert isinstance(back['crs'], LambertConformalSphere)


def test_cf_grid_mapping_attributes():
    assert isinstance(cube_io._crs_from_attrs({'grid_mapping_name': 'latitude_longitude', 'crs_wkt': 'GEOGCRS[...]'}), Geographic)
    lcc = cube_io._crs_from_attrs({'grid_mapping_name': b'lambert_conformal_conic', 'standard_parallel': np.array([38.5, 38.5]),
                                   'latitude_of_projection_origin': np.array([38.5]), 'longitude_of_central_meridian': np.array([262.5]),
                                   'semi_major_axis': np.array([6371229.0]), 'semi_minor_axis': np.array([6371229.0]),
                                   'false_easting': np.array([0.0]), 'false_northing': np.array([0.0])})
    assert isinstance(lcc, LambertConformalSphere) and lcc.args['R'] == 6371229.0
    with pytest.raises(NotImplementedError):   # ellipsoidal Lambert: not the HRRR form
        cube_io._crs_from_attrs({'grid_mapping_name': 'lambert_conformal_conic', 'standard_parallel': [33.0, 45.0], 'latitude_of_projection_origin': 40.0,
                                 'longitude_of_central_meridian': -97.0, 'semi_major_axis': 6378137.0, 'semi_minor_axis': 6356752.3})
    with pytest.raises(NotImplementedError):
        cube_io._crs_from_attrs({'grid_mapping_name': 'polar_stereographic'})
