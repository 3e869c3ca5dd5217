"""Generate tests/golden/era5_slant_ref.npz + orbit_S1B_20200130_sv.txt + era5_gnss_ref.npz -- run HERE (CPU container), never on the GPU box.

    python tests/golden/make_golden_era5_slant.py

The reference's two end-to-end goldens of the slant path on its own ERA-5 fixture (test/test_slant.py):

* ``test_slant_proj``  (test/test_slant.py:49)  2.333865144 m at (33.4, -117.8, 0): a *projected* LOS on a cube AOI -- which
  delay.py:147-160 answers with the zenith totals through ``_build_cube``;
* ``test_ray_tracing`` (test/test_slant.py:99)  2.97711681 m at the same point: ``Raytracing(orbit file)`` through
  ``_build_cube_ray``.

What runs here is the REFERENCE'S OWN Python (oracle/refpy.py): ``llreader.BoundingBox`` + ``add_buffer`` +
``set_output_xygrid`` (cli/raider.py:257-260; the default 2000 m cube spacing, constants.py:22), ``losreader.Raytracing`` with
``get_orbit`` on the reference's precise-orbit file, ``delay._build_cube`` / ``delay._build_cube_ray``.  isce3 is the stand-in
of oracle/refpy.py (= oracle/orbit.py's restatement of isce3's Hermite interpolation and zero-Doppler Newton solve) and the
NetCDF-4 cube is read by raider_b200/hdf5_lite.py: BOTH GOLDENS ARE REPRODUCED TO THE 7 DECIMALS THE REFERENCE ASSERTS, which
is what pins oracle/orbit.py and the HDF5 reader to the reference's published numbers.  The fixture carries the reference
functions' full output cubes (4 x 90 x 90), the ERA-5 cube and the +-600 s of state vectors ``get_sv`` keeps, so the GPU box
can compare the CUDA path with them (tests/test_gpu_round2.py).
"""
from __future__ import annotations

import datetime as dt
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent
REFERENCE = Path('/root/reference')
WM = REFERENCE / 'test' / 'weather_files' / 'ERA-5_2020_01_30_T13_52_45_32N_35N_120W_115W.nc'
EOF = REFERENCE / 'test' / 'orbit_files' / 'S1B_OPER_AUX_POEORB_OPOD_20210317T025713_V20200129T225942_20200131T005942.EOF'
TIME = dt.datetime(2020, 1, 30, 13, 52, 45)
BBOX = [33, 34, -118.25, -117.25]            # test/test_slant.py:23,67
HEIGHTS = [0, 100, 500, 1000]                # test/test_slant.py:29,75
GOLD_POINT = (33.4, -117.8, 0.0)
GOLD_STD, GOLD_RAY = 2.333865144, 2.97711681  # test/test_slant.py:49,99


def reference_setup():
    """(ref namespace, cube dict, aoi) exactly as calcDelays prepares them (cli/raider.py:250-260)."""
    from oracle import refpy
    from raider_b200.cube_io import load_cube
    ref = refpy.load()
    llreader = importlib.import_module('RAiDER.llreader')
    cube = load_cube(WM)
    aoi = llreader.BoundingBox(list(BBOX), cube_spacing_in_m=2000.0)   # cli/types.py:173 default
    aoi.add_buffer(0.25)                                               # ERA5().getLLRes(): models/ecmwf.py:32-33
    aoi.set_output_xygrid(4326)
    return ref, cube, aoi


def gold_index(aoi):
    """xarray's .sel(method='nearest') of test/test_slant.py:45-46."""
    return int(np.argmin(np.abs(aoi.ypts - GOLD_POINT[0]))), int(np.argmin(np.abs(aoi.xpts - GOLD_POINT[1])))


def reference_ztd(ref, cube, aoi, zpts):
    crs = ref.CRS(4326)
    ifs = ref.delayFcns.getInterpolators(ref.dataset(cube), 'total')
    return ref.delay._build_cube(aoi.xpts, aoi.ypts, zpts, crs, crs, ifs)


def reference_ray(ref, cube, aoi, zpts):
    crs = ref.CRS(4326)
    los = ref.losreader.Raytracing(str(EOF), time=TIME)
    ifs = ref.delayFcns.getInterpolators(ref.dataset(cube), 'pointwise')
    toa = cube['z'].max() - 1                                          # delay.py:84-87: zref=None -> top of the model
    out = ref.delay._build_cube_ray(aoi.xpts, aoi.ypts, zpts, los, crs, crs, ifs, MAX_TROPO_HEIGHT=toa)
    return out, los, toa


def main() -> None:
    ref, cube, aoi = reference_setup()
    zpts = np.array(HEIGHTS, dtype=np.float64)
    iy, ix = gold_index(aoi)
    zw, zh = reference_ztd(ref, cube, aoi, zpts)
    np.testing.assert_almost_equal(GOLD_STD, (zw + zh)[0, iy, ix])
    (rw, rh), los, toa = reference_ray(ref, cube, aoi, zpts)
    np.testing.assert_almost_equal(GOLD_RAY, (rw + rh)[0, iy, ix])
    orb = los._orbit
    # the state vectors get_sv kept (losreader.py:368-370), as the 7-column text file read_txt_file accepts (losreader.py:423-475)
    epoch = orb.reference_epoch.t
    with open(OUT / 'orbit_S1B_20200130_sv.txt', 'w') as f:
        for t, p, v in zip(orb.time, orb.position, orb.velocity):
            stamp = (epoch + dt.timedelta(seconds=float(t))).isoformat()
            f.write(stamp + ' ' + ' '.join(repr(float(c)) for c in (*p, *v)) + '\n')
    np.savez_compressed(
        OUT / 'era5_slant_ref.npz',
        x=cube['x'], y=cube['y'], z=cube['z'], wet=cube['wet'], hydro=cube['hydro'],
        wet_total=cube['wet_total'], hydro_total=cube['hydro_total'],
        xpts=aoi.xpts, ypts=aoi.ypts, zpts=zpts, zref=np.float64(toa), gold_index=np.array([iy, ix]),
        gold_std=np.float64(GOLD_STD), gold_ray=np.float64(GOLD_RAY),
        ref_ztd_wet=zw, ref_ztd_hydro=zh, ref_ray_wet=rw, ref_ray_hydro=rh,
        time=np.array(TIME.isoformat()))
    print('std', (zw + zh)[0, iy, ix], 'ray', (rw + rh)[0, iy, ix], 'written', OUT / 'era5_slant_ref.npz')


STATIONS = REFERENCE / 'test' / 'scenario_6' / 'stations.csv'
GOLD_GNSS = ('TORP', 2.34514)                 # test/test_intersect.py:104 (4 decimals)


def reference_station_ztd():
    """test/test_intersect.py::test_gnss_intersect on the same ERA-5 file: a station AOI (llreader.StationFile) is answered by a ZTD
    cube on the AOI's own grid at the model's z levels (delay.py:78-96, 147-160) that is then interpolated at the stations
    (delay.py:104-121).  All of it is the reference's own Python (pandas reads the CSV)."""
    from oracle import refpy
    from raider_b200.cube_io import load_cube
    ref = refpy.load()
    llreader = importlib.import_module('RAiDER.llreader')
    cube = load_cube(WM)
    aoi = llreader.StationFile(str(STATIONS), cube_spacing_in_m=2000.0)
    aoi.add_buffer(0.25)
    aoi.set_output_xygrid(4326)
    crs = ref.CRS(4326)
    zpts = cube['z']                                                  # height_levels = wm_levels (delay.py:80-84)
    zw, zh = ref.delay._build_cube(aoi.xpts, aoi.ypts, zpts, crs, crs, ref.delayFcns.getInterpolators(ref.dataset(cube), 'total'))
    lats, lons = aoi.readLL()
    hgts = aoi.readZ()
    pnts = ref.delay.transformPoints(lats, lons, hgts, crs, crs)
    if_w, if_h = ref.delayFcns.getInterpolators(ref.dataset(dict(x=aoi.xpts, y=aoi.ypts, z=zpts, wet=zw, hydro=zh)), 'ztd')
    return aoi, (lats, lons, hgts), (zw, zh), (if_w(pnts), if_h(pnts))


def main_stations() -> None:
    import pandas as pd
    aoi, (lats, lons, hgts), _, (wd, hd) = reference_station_ztd()
    ids = list(pd.read_csv(STATIONS)['ID'])
    np.testing.assert_almost_equal((wd + hd)[ids.index(GOLD_GNSS[0])], GOLD_GNSS[1], decimal=4)
    np.savez_compressed(OUT / 'era5_gnss_ref.npz', xpts=aoi.xpts, ypts=aoi.ypts, lats=lats, lons=lons, hgts=hgts, ref_wet=wd, ref_hydro=hd,
                        gold_index=np.int64(ids.index(GOLD_GNSS[0])), gold_total=np.float64(GOLD_GNSS[1]))
    print('stations', dict(zip(ids, wd + hd)), 'written', OUT / 'era5_gnss_ref.npz')


if __name__ == '__main__':
    main()
    main_stations()
