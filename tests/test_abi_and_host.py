"""CPU-side tests: the C-ABI library builds, loads and exports exactly what include/raider_b200.h declares; the product
fails loudly without a GPU (no fallback); host logic (CRS parsing, cube I/O, AOI grids, Npts rule, layer counting)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _header_functions():
    text = (ROOT / 'include' / 'raider_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(rdr_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol(lib):
    from raider_b200 import _lib
    declared = _header_functions()
    assert len(declared) >= 20
    assert sorted(_lib.SIGNATURES) == declared, 'ctypes SIGNATURES and include/raider_b200.h disagree'
    raw = C.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(raw, name), f'{name} is declared in the header but not exported by the .so'
    assert lib.rdr_abi_version() == 3


def test_no_extra_exports(lib):
    """Only the rdr_* C symbols are visible (-fvisibility=hidden): no C++/CUDA internals leak through the boundary."""
    import subprocess
    from raider_b200 import _lib
    out = subprocess.run(['nm', '-D', '--defined-only', str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    names = [ln.split()[-1] for ln in out.splitlines() if ' T ' in ln]
    extra = [n for n in names if not n.startswith('rdr_') and n not in ('_init', '_fini')]
    assert not extra, extra


def test_sass_is_sm100a_and_uses_tma_bulk():
    """The shipped cubin targets sm_100a and the streaming sampler really goes through the TMA engine (UBLKCP) + mbarriers."""
    import shutil
    import subprocess
    from raider_b200 import _lib, build
    build.build()
    if not shutil.which('cuobjdump'):
        pytest.skip('cuobjdump not on PATH')
    lst = subprocess.run(['cuobjdump', '-lelf', str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert 'sm_100a' in lst
    sass = subprocess.run(['cuobjdump', '-sass', str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert 'UBLKCP' in sass and 'SYNCS' in sass
    assert 'MUFU.RSQ64H' in sass and 'MUFU.RCP64H' in sass


def test_fails_loudly_without_gpu(lib):
    """On a box without a CUDA device every product entry point raises; nothing silently computes on the CPU."""
    if lib.rdr_device_count() > 0:
        pytest.skip('a GPU is visible here')
    from raider_b200 import _lib, synthetic as syn
    from raider_b200.delayFcns import getInterpolators
    from raider_b200.interpolate import interpolate
    from raider_b200.makePoints import makePoints0D
    with pytest.raises(_lib.RaiderB200Error, match='no CPU fallback'):
        _lib.Handle(0)
    with pytest.raises(RuntimeError):
        getInterpolators(syn.config_c1()['cube'])
    with pytest.raises(RuntimeError):
        interpolate((np.array([0.0, 1.0]),), np.array([0.0, 1.0]), np.array([[0.5]]))
    with pytest.raises(RuntimeError):
        makePoints0D(10.0, np.zeros(3), np.ones(3), 5.0)


def test_missing_library_is_an_import_error(monkeypatch, tmp_path):
    from raider_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setenv('RAIDER_B200_LIB', str(tmp_path / 'nope.so'))
    with pytest.raises(ImportError, match='no CPU fallback'):
        _lib.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under raider_b200/ may import it."""
    for p in (ROOT / 'raider_b200').rglob('*.py'):
        src = p.read_text()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), p


def test_argument_validation_needs_no_gpu(lib):
    """Shape / dtype errors are raised by the shims before anything touches the device (module.cpp:36-63 conventions)."""
    from raider_b200.interpolate import interpolate, interpolate_along_axis
    from raider_b200.makePoints import makePoints1D
    with pytest.raises(TypeError):
        interpolate((np.zeros(10), np.zeros(5)), np.zeros(1), np.zeros(1))
    with pytest.raises(TypeError):
        interpolate((np.zeros(3),), np.zeros(3), np.zeros(3))          # interp_points must be (N, ndim)
    with pytest.raises(TypeError):
        interpolate_along_axis(np.array(0), np.array(0), np.array(0))
    with pytest.raises(TypeError):
        interpolate_along_axis(np.zeros(1), np.zeros(1), np.zeros(1), axis=-2)
    with pytest.raises(RuntimeError):
        interpolate_along_axis(np.zeros((2, 2)), np.zeros((2, 2)), np.zeros((2, 2)), axis=0)
    with pytest.raises(ValueError):
        makePoints1D(10.0, np.zeros((2, 3), dtype=np.float32), np.zeros((2, 3)), 5.0)
    with pytest.raises(ZeroDivisionError):
        makePoints1D(10.0, np.zeros((2, 3)), np.zeros((2, 3)), 0.0)


def test_make_points_count_rule(lib, golden):
    """Npts rule (makePoints.pyx:130-134 as Cython compiles it) is host code in the library: checked against the compiled reference."""
    from oracle import interp as ointerp
    g = golden('makepoints')
    n = C.c_int64(0)
    for L, s, want in g['counts']:
        assert lib.rdr_make_points_count(float(L), float(s), C.byref(n)) == 0
        assert n.value == int(want) == ointerp.make_npts(L, s)
    rng = np.random.default_rng(0)
    for L, s in zip(rng.uniform(0.1, 5e4, 300), rng.uniform(0.01, 500, 300)):
        lib.rdr_make_points_count(float(L), float(s), C.byref(n))
        assert n.value == ointerp.make_npts(L, s)


def test_layer_count_matches_oracle_plan(lib):
    """rdr_build_ray's count query runs the scalar layer decisions of losreader.py:785-809 on the host."""
    from oracle import raytrace as rt
    from raider_b200 import synthetic as syn
    for zs in (syn.z_levels(37), syn.z_levels_table('ml145'), syn.z_levels_table('hrrr57')):
        for ht, zref in [(0.0, zs[-1] - 1), (1500.0, 26000.0), (-500.0, 9e4), (zs[-1] - 0.5, zs[-1] - 1), (9.5, 12000.0), (zs[5], zs[5] + 0.5)]:
            K = C.c_int64(-1)
            rc = lib.rdr_build_ray(zs.ctypes.data, zs.size, float(ht), None, None, 0, float(zref), C.byref(K), None, None, None, 0)
            want = len(rt.layer_plan(zs, ht, zref))
            assert K.value == want and rc == (0 if want else 4)


def test_crs_parsing():
    from raider_b200.crs import Geographic, LambertConformalSphere, parse_crs
    assert isinstance(parse_crs(4326), Geographic) and isinstance(parse_crs('EPSG:4326'), Geographic) and isinstance(parse_crs(None), Geographic)
    assert parse_crs(4326) == 4326 and parse_crs(4326) == parse_crs('+proj=longlat +datum=WGS84 +no_defs')
    hrrr = parse_crs('+proj=lcc +lat_1=38.5 +lat_2=38.5 +lat_0=38.5 +lon_0=262.5 +x_0=0 +y_0=0 +a=6371229 +b=6371229 +units=m +no_defs')
    assert isinstance(hrrr, LambertConformalSphere) and hrrr == LambertConformalSphere() and hrrr != parse_crs(4326)
    from oracle.geodesy import LambertConformalSphere as OracleLCC
    assert np.array_equal(hrrr.params(), OracleLCC().params())
    lon, lat = np.meshgrid(np.linspace(-120, -75, 7), np.linspace(25, 50, 5))
    x, y = hrrr.from_ll(lon, lat)
    ox, oy = OracleLCC().forward(lon, lat)
    assert np.array_equal(x, ox) and np.array_equal(y, oy)
    lo, la, _ = hrrr.to_llh(x, y, 0 * x)
    assert np.allclose(np.mod(lo, 360), np.mod(lon, 360), atol=1e-10) and np.allclose(la, lat, atol=1e-10)
    with pytest.raises(NotImplementedError):
        parse_crs(32611)
    with pytest.raises(NotImplementedError):
        parse_crs('+proj=lcc +lat_1=33 +lat_2=45 +lat_0=40 +lon_0=-97 +ellps=GRS80')

    class FakePyproj:  # duck type of pyproj.CRS
        def to_epsg(self):
            return 4326
    assert isinstance(parse_crs(FakePyproj()), Geographic)


def test_cube_io_roundtrip(tmp_path):
    from raider_b200 import synthetic as syn
    from raider_b200.crs import Geographic
    from raider_b200.cube_io import load_cube, write_cube
    cube = syn.config_c1()['cube']
    for name in ('wm.nc', 'wm.npz'):
        back = load_cube(write_cube(tmp_path / name, cube))
        assert isinstance(back['crs'], Geographic)
        for k in ('x', 'y', 'z', 'wet', 'hydro', 'wet_total', 'hydro_total'):
            assert np.array_equal(np.asarray(back[k], dtype=cube[k].dtype), cube[k]), (name, k)
    assert load_cube(cube)['crs'] is None
    # a truncated / corrupt HDF5 file is refused by the format reader (tests/test_cube_io_hdf5.py reads the reference's real files)
    (tmp_path / 'hdf.nc').write_bytes(b'\x89HDF\r\n\x1a\n' + b'0' * 64)
    with pytest.raises(ValueError):
        load_cube(tmp_path / 'hdf.nc')


def test_synthetic_totals_match_reference_trapz():
    """cumulative_total == the loop of weatherModel.py:398-401 (np.trapz from each level to the top)."""
    from oracle import raytrace as rt
    from raider_b200 import synthetic as syn
    cube = syn.config_c1()['cube']
    want = rt.cumulative_ztd(cube['wet'], cube['z'])
    assert np.allclose(syn.cumulative_total(cube['wet'], cube['z']), want, rtol=1e-13, atol=1e-18)
    assert np.all(np.diff(syn.z_levels_table('ml145')) > 0) and syn.z_levels_table('ml145').size == 145 and syn.z_levels_table('hrrr57').size == 57


def test_aoi_grid_and_results_dataset(tmp_path):
    import datetime as dt
    from raider_b200.cube_io import load_cube
    from raider_b200.delay import transformPoints, writeResultsToXarray
    from raider_b200.llreader import BoundingBox, Points, is_cube_aoi
    aoi = BoundingBox([33.0, 34.0, -118.0, -117.0], spacing=0.25)
    assert np.allclose(aoi.xpts, [-118, -117.75, -117.5, -117.25, -117.0]) and aoi.ypts[0] == 34.0 and np.all(np.diff(aoi.ypts) < 0)  # llreader.py:190-191
    assert is_cube_aoi(aoi) and not is_cube_aoi(Points([33.5], [-117.5], [10.0]))
    ds = writeResultsToXarray(dt.datetime(2020, 1, 30), aoi.xpts, aoi.ypts, np.array([0.0, 100.0]), 4326, np.zeros((2, 5, 5)), np.ones((2, 5, 5)),
                              'ERA5_x.nc', 'slant - raytracing')
    assert ds['wet'].attrs['units'] == 'm' and ds['hydro'].attrs['description'] == 'hydrostatic slant - raytracing delay'
    assert ds.attrs['source'] == 'ERA5_x.nc' and ds['y'].attrs['units'] == 'degrees_north'
    if hasattr(ds, 'to_netcdf') and type(ds).__name__ == 'SimpleDataset':
        ds.to_netcdf(tmp_path / 'out.nc')
        assert load_cube(tmp_path / 'out.nc')['hydro'].shape == (2, 5, 5)
    pts = transformPoints(np.array([33.5]), np.array([-117.5]), np.array([12.0]), 4326, 4326)
    assert pts.shape == (1, 3) and np.array_equal(pts[0], [33.5, -117.5, 12.0])


def test_orbit_file_readers_and_windowing():
    """Host side of the orbit LOS (mirrors test/test_losreader.py:95-175 on the same fixture values)."""
    import datetime as dt
    from conftest import GOLDEN
    from raider_b200.losreader import Orbit, cut_times, filter_ESA_orbit_file, get_orbit, get_sv, read_ESA_Orbit_file, read_txt_file
    txt, eof = GOLDEN / 'orbit_S1_sv.txt', GOLDEN / 'orbit_S1_example.EOF'
    a, b = read_txt_file(txt), read_ESA_Orbit_file(eof)
    assert len(a) == 7 and a[0][0] == dt.datetime(2018, 11, 12, 23, 0, 2) and a[0][-1] == dt.datetime(2018, 11, 12, 23, 1, 12)
    assert all(np.allclose(x, y) for x, y in zip(a[1:], b[1:])) and list(a[0]) == list(b[0])
    assert np.isclose(a[1][0], -2064965.285362) and np.isclose(a[6][-1], -7235.952940)
    with pytest.raises(ValueError):
        read_txt_file(eof)
    t = a[0]
    assert all(cut_times(t, t[0], pad=3600 * 3)) and sum(cut_times(t, t[0], pad=5)) == 1
    assert np.sum(cut_times(t, t[4], pad=15)) == 3 and np.sum(cut_times(t, t[0], pad=400)) == len(t)
    sv = get_sv(str(txt), dt.datetime(2018, 11, 12, 23, 0, 32), pad=25)
    assert sv[0].size == 5
    with pytest.raises(ValueError):
        get_sv(str(GOLDEN / 'make_golden.py'), dt.datetime(2018, 11, 12, 23, 0, 32), pad=25)
    name = 'S1A_OPER_AUX_POEORB_OPOD_20181203T120749_V20181112T225942_20181114T005942.EOF'
    assert filter_ESA_orbit_file(name, dt.datetime(2018, 11, 13, 1, 0, 0)) and not filter_ESA_orbit_file(name, dt.datetime(2018, 11, 15))
    orb = get_orbit(str(txt), dt.datetime(2018, 11, 12, 23, 0, 30), 600)
    assert orb.size == 8 and np.array_equal(orb.time, 10.0 * np.arange(8)) and orb.packed().shape == (1 + 8 * 7,)
    assert orb.packed()[0] == 8.0 and orb.packed()[1 + 7 + 1] == a[1][1]
    shuffled = Orbit(a[0][[2, 0, 1, 3, 3, 4]], np.stack(a[1:4], -1)[[2, 0, 1, 3, 3, 4]], np.stack(a[4:7], -1)[[2, 0, 1, 3, 3, 4]])
    assert np.array_equal(shuffled.time, [0.0, 10.0, 20.0, 30.0, 40.0])
    with pytest.raises(ValueError):
        Orbit(a[0][:3], np.stack(a[1:4], -1)[:3], np.stack(a[4:7], -1)[:3])


def test_symmetric_maps_peer_addresses():
    """Address arithmetic of the fused gather (raider_b200.dist.SymmetricMaps.peer_ptrs): row block [r0, r1) of height slice hh
    inside every other rank's (2, nz, ny, nx) float64 maps -- no GPU needed for the arithmetic."""
    from types import SimpleNamespace
    from raider_b200.dist import SymmetricMaps
    sm = object.__new__(SymmetricMaps)
    sm.comm = SimpleNamespace(rank=1, world=3)
    sm.nz, sm.ny, sm.nx = 2, 10, 7
    sm.base = [1000, 5000, 9000]
    wet, hydro = sm.peer_ptrs(1, 4, 6)
    plane = 10 * 7 * 8
    assert wet == [1000 + (1 * plane + 4 * 7 * 8), 9000 + (1 * plane + 4 * 7 * 8)]                 # ranks 0 and 2, height slice 1, row 4
    assert hydro == [b + 2 * plane for b in wet]                                                    # the hydro maps follow the nz wet slices
    wet_all, hydro_all = sm.peer_ptrs(0, 0, 10, include_self=True)
    assert wet_all == [1000, 5000, 9000] and hydro_all == [1000 + 2 * plane, 5000 + 2 * plane, 9000 + 2 * plane]


@pytest.mark.timeout(600)
def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU port of the reference loops on all host cores): one JSON line with the contract's keys,
    the same metric / unit / workload as the GPU arm, no GPU launches."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    res = subprocess.run([sys.executable, str(root / 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'], capture_output=True,
                         text=True, timeout=550, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'rays/s' and line['higher_is_better'] is True and line['vs_baseline'] is None
    assert line['value'] > 0 and line['gpu_launches'] == 0 and line['dtype'] == 'f64' and line['data'] == 'synthetic'
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1 and line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    sys.path.insert(0, str(root))
    import bench
    assert line['metric'] == bench.METRIC and line['config']['workload'].startswith('C2 slant delay: 2000x2000 rays per GPU')
