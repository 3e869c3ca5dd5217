"""Tropospheric delay calculation on the B200 (reference: tools/RAiDER/delay.py).

Same functions, same signatures: ``tropo_delay``, ``_get_delays_on_cube``, ``_build_cube``, ``_build_cube_ray``,
``writeResultsToXarray``, ``transformPoints``.  The Python bodies only unpack arguments and loop over output heights;
every sample, transform and sum happens in libraider_b200.so (K0/K3 for ray tracing, K2 for zenith/projected and for
the point-mode re-interpolation).
"""
from __future__ import annotations

import datetime as dt
import logging
import os
from typing import Optional, Union

import numpy as np

from . import _lib
from .constants import _ZREF
from .crs import Geographic, parse_crs
from .cube_io import load_cube
from .delayFcns import DeviceInterpolator, as_device_cube, getInterpolators
from .engine import DeviceCube, TraceInfo, los_device_spec
from .llreader import is_cube_aoi
from .utilFcns import lla2ecef

logger = logging.getLogger('RAiDER')

# cross-GPU hooks installed by raider_b200.dist.enable(); None = single GPU
_reduce_hooks = None


###############################################################################
def tropo_delay(
    datetime: dt.datetime,
    weather_model_file,
    aoi,
    los,
    height_levels: Optional[list] = None,
    out_proj: Union[int, str] = 4326,
    zref: Optional[np.float64] = None,
):
    """Calculate integrated delays on query points (delay.py:35-130).

    1. Zenith delays (ZTD)  2. Zenith delays projected to the line-of-sight  3. Slant delays by ray tracing.
    Returns ``(Dataset, None)`` for cube AOIs, else ``(wetDelay, hydroDelay)`` arrays at the query points.
    """
    crs = parse_crs(out_proj)

    ds_in = load_cube(weather_model_file)
    wm_proj = ds_in.get('crs')
    if wm_proj is None:
        logger.warning("WARNING: I can't find a CRS in the weather model file, so I will assume you are using WGS84")
        wm_proj = parse_crs(4326)

    wm_levels = np.asarray(ds_in['z'])
    toa = wm_levels.max() - 1

    if height_levels is None:
        if type(aoi).__name__ == 'Geocube':
            height_levels = aoi.readZ()
        else:
            height_levels = wm_levels

    if zref is None:
        zref = toa

    if zref > toa:
        zref = toa
        logger.warning(f'Requested integration height (zref) is higher than top of weather model. Forcing to top ({toa}).')

    ds = _get_delays_on_cube(datetime, ds_in, wm_proj, aoi, height_levels, los, crs, zref)

    if is_cube_aoi(aoi):
        return ds, None

    pnt_proj = parse_crs(4326)
    lats, lons = aoi.readLL()
    hgts = aoi.readZ()
    pnts = transformPoints(lats, lons, hgts, pnt_proj, crs)

    try:
        wetDelay, hydroDelay = _interp_delay_cube(ds, pnts)
    except RuntimeError:
        raise RuntimeError(f'Failed to get weather model {weather_model_file} interpolators.')

    # return the delays (ZTD or STD)
    if los.is_Projected():
        los.setTime(datetime)
        los.setPoints(lats, lons, hgts)
        wetDelay = los(wetDelay)
        hydroDelay = los(hydroDelay)

    return wetDelay, hydroDelay


def _interp_delay_cube(ds, pnts):
    """delay.py:116-121: ``getInterpolators(ds, 'ztd')`` then evaluate at the query points.

    The delay cube is float64; :class:`DeviceCube` stages float64 fields as hi + lo float32 parts (value == hi + lo to ~2^-48)
    and samples both -- linear interpolation is linear in the values.
    """
    x, y, z = (np.asarray(ds.variables[k][:], dtype=np.float64) for k in ('x', 'y', 'z'))
    cube = DeviceCube(y, x, z, np.asarray(ds.variables['wet'][:], dtype=np.float64), np.asarray(ds.variables['hydro'][:], dtype=np.float64),
                      layout=_lib.LAYOUT_ZYX)
    return cube.sample(pnts)


def _get_delays_on_cube(datetime, weather_model_file, wm_proj, aoi, heights, los, crs, zref, nproc=1):
    """Raider cube generation function (delay.py:133-193)."""
    zpts = np.array(heights)
    ds_in = load_cube(weather_model_file)
    wm_proj = parse_crs(wm_proj)
    crs = parse_crs(crs)

    try:
        aoi.xpts
    except AttributeError:
        x_spacing = np.diff(np.asarray(ds_in['x'])).mean()
        y_spacing = np.diff(np.asarray(ds_in['y'])).mean()
        aoi.set_output_spacing(ll_res=np.min([x_spacing, y_spacing]))
        aoi.set_output_xygrid(crs if not isinstance(crs, Geographic) else 4326)

    if los.is_Zenith() or los.is_Projected():
        out_type = ['zenith' if los.is_Zenith() else 'slant - projected'][0]
        ds_in['crs'] = wm_proj
        ifWet, ifHydro = getInterpolators(ds_in, 'total')
        wetDelay, hydroDelay = _build_cube(aoi.xpts, aoi.ypts, zpts, wm_proj, crs, [ifWet, ifHydro])
    else:
        out_type = 'slant - raytracing'
        ds_in['crs'] = wm_proj
        ifWet, ifHydro = getInterpolators(ds_in, kind='pointwise', shared=(nproc > 1))
        if nproc == 1:
            wetDelay, hydroDelay = _build_cube_ray(aoi.xpts, aoi.ypts, zpts, los, wm_proj, crs, [ifWet, ifHydro], MAX_TROPO_HEIGHT=zref)
        else:
            raise NotImplementedError  # as in the reference (delay.py:178-185): parallelism lives in raider_b200.dist

    if np.isnan(wetDelay).any() or np.isnan(hydroDelay).any():
        logger.critical('There are missing delay values. Check your inputs.')

    name = weather_model_file if isinstance(weather_model_file, (str, os.PathLike)) else 'in-memory cube'
    ds = writeResultsToXarray(datetime, aoi.xpts, aoi.ypts, zpts, crs, wetDelay, hydroDelay, name, out_type)
    return ds


def _build_cube(xpts, ypts, zpts, model_crs, pts_crs, interpolators):
    """Iterate over interpolators and build a cube using Zenith (delay.py:196-216)."""
    xpts, ypts, zpts = (np.asarray(a, dtype=np.float64) for a in (xpts, ypts, zpts))
    model_crs, pts_crs = parse_crs(model_crs), parse_crs(pts_crs)
    cube = as_device_cube(interpolators, crs=model_crs) if len(interpolators) == 2 else None
    if cube is not None and model_crs == pts_crs:
        return list(cube.sample_grid_levels(xpts, ypts, zpts))   # the whole height loop in one launch
    outputArrs = [np.zeros((zpts.size, ypts.size, xpts.size)) for mm in range(len(interpolators))]

    for ii, ht in enumerate(zpts):
        if model_crs != pts_crs:
            xx, yy = np.meshgrid(xpts, ypts)
            pts = transformPoints(yy, xx, np.full(yy.shape, ht), pts_crs, model_crs)
            if cube is not None:
                vals = cube.sample(pts)
            else:
                vals = [intp(pts) for intp in interpolators]
        elif cube is not None:
            vals = cube.sample_grid(xpts, ypts, ht)
        else:
            xx, yy = np.meshgrid(xpts, ypts)
            pts = np.stack([yy, xx, np.full(yy.shape, ht)], axis=-1)
            vals = [intp(pts) for intp in interpolators]
        for mm in range(len(interpolators)):
            outputArrs[mm][ii, ...] = vals[mm]

    return outputArrs


def _build_cube_ray(
    xpts,
    ypts,
    zpts,
    los,
    model_crs,
    pts_crs,
    interpolators,
    outputArrs=None,
    MAX_SEGMENT_LENGTH=1000.0,
    MAX_TROPO_HEIGHT=_ZREF,
    _out_device=None,
    _out_arrays=None,
    _peers=None,
    _exchange=None,
    _on_skip=None,
):
    """Iterate over interpolators and build a cube using raytracing (delay.py:219-326).

    MAX_TROPO_HEIGHT should not extend above the top of the weather model.  ``outputArrs`` (two (nz, ny, nx) float64
    arrays) is accumulated into in place when given (:245-248,:323), else a new list is returned.
    """
    xpts, ypts, zpts = (np.asarray(a, dtype=np.float64) for a in (xpts, ypts, np.atleast_1d(zpts)))
    if len(interpolators) != 2:
        raise TypeError('the device path integrates the (ifWet, ifHydro) pair in one pass: pass both interpolators')
    model_crs, pts_crs = parse_crs(model_crs), parse_crs(pts_crs)
    cube = as_device_cube(interpolators, crs=model_crs)
    if cube.crs != model_crs:
        raise ValueError(f'model_crs {model_crs} does not match the CRS the cube was staged with ({cube.crs})')

    ny, nx = ypts.size, xpts.size
    output_created_here = False
    if outputArrs is None:
        # np.zeros((nz, ny, nx)) in the reference (:248); here each slice is written straight from the device (0 + x == x),
        # so the arrays start uninitialised and only skipped slices are zero-filled
        output_created_here = True
        if _out_arrays is not None:   # raider_b200.dist: this rank's row block inside the symmetric (peer-mapped) full maps
            outputArrs = list(_out_arrays)
        elif _out_device is not None:   # raider_b200.dist: keep the (nz, ny, nx) maps in HBM for the all-gather
            import torch
            outputArrs = [torch.empty((zpts.size, ny, nx), dtype=torch.float64, device=_out_device) for mm in range(2)]
        else:
            outputArrs = [_lib.pinned_empty((zpts.size, ny, nx)) for mm in range(2)]   # page-locked: the kernel writes them directly
    else:
        # scratch for one height in page-locked memory: the kernel writes it directly (no staged device-to-host copy), the
        # accumulation into the caller's arrays (:323) stays on the host
        wet = _lib.pinned_empty((ny, nx))
        hydro = _lib.pinned_empty((ny, nx))

    spec = los_device_spec(los, ny, nx)
    geographic_pts = isinstance(pts_crs, Geographic)
    hooks = _reduce_hooks or (None, None)
    cube.last_info = []

    for hh, ht in enumerate(zpts):
        logger.info(f'Processing slice {hh+1} / {len(zpts)}: {ht}')
        if output_created_here:
            wet, hydro = outputArrs[0][hh], outputArrs[1][hh]
        # Step 1 + 2: ground points and look vectors.  Regular geographic rasters with a device-generated LOS never
        # leave the GPU; anything else goes through the host geometry layer exactly like the reference (:262-270).
        if geographic_pts and spec is not None:
            geom = (_lib.GEOM_GRID, xpts, ypts)
            los_kind, los_payload = spec
        else:
            xx, yy = np.meshgrid(xpts, ypts)
            llh = [xx, yy, np.full(yy.shape, ht)] if geographic_pts else list(pts_crs.to_llh(xx, yy, np.full(yy.shape, ht)))
            if spec is not None:
                los_kind, los_payload = spec
            else:
                xyz = np.stack(lla2ecef(llh[1], llh[0], llh[2]), axis=-1)
                LOS = np.asarray(los.getLookVectors(ht, llh, xyz, yy), dtype=np.float64)
                if LOS.shape != (ny, nx, 3):
                    raise ValueError(f'getLookVectors returned shape {LOS.shape}, expected {(ny, nx, 3)}')
                los_kind, los_payload = _lib.LOS_ARRAY, np.ascontiguousarray(LOS.reshape(-1, 3))
            geom = (_lib.GEOM_POINTS, np.ascontiguousarray(llh[0], dtype=np.float64).ravel(),
                    np.ascontiguousarray(llh[1], dtype=np.float64).ravel())

        # Steps 3..: layers, nParts, sub-steps, sampling, trapezoid -- all on the device
        try:
            info = cube.trace(geom[0], geom[1], geom[2], ny, nx, los_kind, los_payload, ht, MAX_TROPO_HEIGHT, MAX_SEGMENT_LENGTH,
                              wet, hydro, reduce_max=hooks[0], reduce_sum=hooks[1],
                              peers_fn=(lambda r0, r1, hh=hh: _peers(hh, r0, r1)) if (_peers is not None and output_created_here) else None,
                              exchange=_exchange)
        except _lib.NoLayersError:
            # if the top most height layer doesnt contribute to the integral, skip it (:276-277)
            if ht == zpts[-1]:
                cube.last_info.append(TraceInfo(ht=float(ht), skipped=True))
                if output_created_here:
                    wet[...] = 0.0
                    hydro[...] = 0.0
                if _on_skip is not None:   # raider_b200.dist: the rows of the other ranks in this rank's full maps are zero as well
                    _on_skip(hh)
                continue
            # the reference evaluates np.isnan(None) here and dies with a TypeError (:279); say why instead
            raise TypeError(f'no weather-model layer contributes between height {ht} and MAX_TROPO_HEIGHT={MAX_TROPO_HEIGHT}')
        cube.last_info.append(info)
        if not output_created_here:
            outputArrs[0][hh, ...] += wet
            outputArrs[1][hh, ...] += hydro

    if output_created_here:
        return outputArrs


def slant_delay_points(weather_model, lats, lons, hgts, los, zref=None, MAX_SEGMENT_LENGTH=1000.0, second_epoch=None, weights=None):
    """Ray-traced slant delays at explicit points (GNSS stations), each with its own height: BASELINE config C4.

    Every point is traced exactly as the reference traces a one-pixel raster at that height --
    ``_build_cube_ray(xpts=[lon], ypts=[lat], zpts=[h], ...)`` (delay.py:219-326): own layer plan, own ``nParts``, own
    clamps -- but all points go through one kernel (K5, a warp per ray).  ``weather_model``: file / cube dict / interpolator
    pair; ``second_epoch`` + ``weights=(w0, w1)`` fuse the temporal interpolation of cli/raider.py:817-819 at staging.
    ``los``: a Raytracing object (orbit / constant incidence / per-point incidence arrays), ``Zenith``-like ray tracing, or an
    (n, 3) array of ECEF unit vectors.  Returns ``(wet, hydro)`` in metres.
    """
    lats, lons, hgts = (np.asarray(a, dtype=np.float64).ravel() for a in np.broadcast_arrays(lats, lons, hgts))
    if isinstance(weather_model, (list, tuple)) and len(weather_model) == 2:
        cube = as_device_cube(list(weather_model))
    else:
        ds_in = load_cube(weather_model)
        cube = getInterpolators(ds_in, 'pointwise')[0].cube
    if second_epoch is not None:
        w0, w1 = weights
        ds1 = load_cube(second_epoch)
        cube.blend(ds1['wet'], ds1['hydro'], w0, w1)
    toa = float(cube.grid[2].max() - 1)
    zref = toa if zref is None else min(float(zref), toa)
    if isinstance(los, np.ndarray):
        kind, payload = _lib.LOS_ARRAY, np.ascontiguousarray(los, dtype=np.float64).reshape(-1, 3)
    else:
        spec = los_device_spec(los, 1, lats.size)
        inc = getattr(los, '_incidence', None)
        if spec is not None:
            kind, payload = spec
        elif inc is not None:  # per-point incidence / heading arrays -> ENU on the host (trivial), ECEF on the device
            from .losreader import inc_hd_to_enu
            hd = getattr(los, '_heading', None)
            enu = inc_hd_to_enu(np.broadcast_to(np.asarray(inc, dtype=np.float64), lats.shape),
                                np.broadcast_to(np.asarray(0.0 if hd is None else hd, dtype=np.float64), lats.shape))
            kind, payload = _lib.LOS_ENU_ARRAY, np.ascontiguousarray(enu)
        else:
            xyz = np.stack(lla2ecef(lats, lons, hgts), axis=-1)
            vec = np.asarray(los.getLookVectors(hgts, [lons[None], lats[None], hgts[None]], xyz[None], lats[None]), dtype=np.float64)
            kind, payload = _lib.LOS_ARRAY, np.ascontiguousarray(vec.reshape(-1, 3))
    wet, hydro, ns = cube.trace_stations(lons, lats, hgts, kind, payload, zref, MAX_SEGMENT_LENGTH)
    cube.last_station_samples = ns
    return wet, hydro


class _Var:
    def __init__(self, dims, data, attrs=None) -> None:
        self.dims, self.data, self.attrs = tuple(dims), np.asarray(data), dict(attrs or {})

    @property
    def values(self):
        return self.data

    def __getitem__(self, item):
        return self.data[item]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)


class SimpleDataset:
    """Tiny stand-in for the xarray.Dataset that writeResultsToXarray returns (xarray is not installable offline).

    Supports what the delay path and its callers touch: ``ds['wet']``, ``ds.variables[...]``, ``.attrs``,
    ``ds.x / ds.y / ds.z`` and ``to_netcdf(path)`` (NetCDF-3 classic with the reference's variable layout).
    """

    def __init__(self, data_vars, coords, attrs) -> None:
        self.variables = {}
        for k, (dims, data) in coords.items():
            self.variables[k] = _Var(dims, data)
        for k, (dims, data, a) in data_vars.items():
            self.variables[k] = _Var(dims, data, a)
        self.attrs = dict(attrs)

    def __getitem__(self, k):
        return self.variables[k]

    def __setitem__(self, k, v):
        self.variables[k] = v if isinstance(v, _Var) else _Var((), v)

    def __getattr__(self, k):
        try:
            return self.__dict__['variables'][k]
        except KeyError:
            raise AttributeError(k)

    def to_netcdf(self, path):
        from scipy.io import netcdf_file
        with netcdf_file(str(path), 'w', version=2) as nc:
            for k, v in self.attrs.items():
                setattr(nc, k, v)
            for d in ('z', 'y', 'x'):
                nc.createDimension(d, self.variables[d].data.size)
            for k, v in self.variables.items():
                var = nc.createVariable(k, 'i4' if v.data.dtype.kind == 'i' else 'f8', v.dims)
                if v.dims:
                    var[:] = v.data
                else:
                    var.data[()] = v.data
                for ak, av in v.attrs.items():
                    setattr(var, ak, av)
        return path


def writeResultsToXarray(datetime, xpts, ypts, zpts, crs, wetDelay, hydroDelay, weather_model_file, out_type):
    """Pack the delay cube with the reference's variable names and CF attributes (delay.py:329-401).

    Returns an ``xarray.Dataset`` when xarray is importable, else a :class:`SimpleDataset` with the same members.
    """
    crs = parse_crs(crs)
    wet_attrs = {'units': 'm', 'description': f'wet {out_type} delay', 'grid_mapping': 'crs'}
    hydro_attrs = {'units': 'm', 'description': f'hydrostatic {out_type} delay', 'grid_mapping': 'crs'}
    attrs = dict(
        Conventions='CF-1.7',
        title='RAiDER geo cube',
        source=os.path.basename(str(weather_model_file)),
        history=str(dt.datetime.now(tz=dt.timezone.utc)) + ' RAiDER',
        description=f'RAiDER geo cube - {out_type}',
        reference_time=datetime.strftime('%Y%m%dT%H:%M:%S') if hasattr(datetime, 'strftime') else str(datetime),
    )
    degrees = isinstance(crs, Geographic)
    try:
        import xarray as xr
    except ImportError:
        xr = None
    if xr is not None:
        ds = xr.Dataset(
            data_vars=dict(wet=(['z', 'y', 'x'], wetDelay, wet_attrs), hydro=(['z', 'y', 'x'], hydroDelay, hydro_attrs)),
            coords=dict(x=(['x'], xpts), y=(['y'], ypts), z=(['z'], zpts)),
            attrs=attrs,
        )
        ds['crs'] = -2147483647
    else:
        ds = SimpleDataset(
            data_vars=dict(wet=(['z', 'y', 'x'], wetDelay, wet_attrs), hydro=(['z', 'y', 'x'], hydroDelay, hydro_attrs)),
            coords=dict(x=(['x'], xpts), y=(['y'], ypts), z=(['z'], zpts)),
            attrs=attrs,
        )
        ds['crs'] = _Var((), np.int32(-2147483647))
    ds['crs'].attrs['grid_mapping_name'] = 'latitude_longitude' if degrees else 'lambert_conformal_conic'
    ds['z'].attrs.update(axis='Z', units='m', description='height above ellipsoid')
    if degrees:
        ds['y'].attrs.update(units='degrees_north', standard_name='latitude', long_name='latitude')
        ds['x'].attrs.update(units='degrees_east', standard_name='longitude', long_name='longitude')
    else:
        ds['y'].attrs.update(axis='Y', standard_name='projection_y_coordinate',
                             long_name='y-coordinate in projected coordinate system', units='m')
        ds['x'].attrs.update(axis='X', standard_name='projection_x_coordinate',
                             long_name='x-coordinate in projected coordinate system', units='m')
    return ds


def transformPoints(lats, lons, hgts, old_proj, new_proj) -> np.ndarray:
    """Transform lat/lon/hgt points to a new projection; returns (..., 3) in (y, x, z) order (delay.py:404-436)."""
    old_proj, new_proj = parse_crs(old_proj), parse_crs(new_proj)
    lats, lons, hgts = np.broadcast_arrays(np.asarray(lats, dtype=np.float64), np.asarray(lons, dtype=np.float64),
                                           np.asarray(hgts, dtype=np.float64))
    lon, lat, h = old_proj.to_llh(lons, lats, hgts)
    x, y = new_proj.from_ll(lon, lat)
    return np.stack([y, x, h], axis=-1)
