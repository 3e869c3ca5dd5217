"""Reading/writing processed weather-model cubes (the file the delay path consumes).

On-disk contract (reference writer tools/RAiDER/models/weatherModel.py:659-724; readers tools/RAiDER/delay.py:66-78,
tools/RAiDER/delayFcns.py:31-41): variables ``wet, hydro, wet_total, hydro_total`` with dims (z, y, x), coordinate
variables ``x, y, z`` and a ``proj`` variable carrying ``crs_wkt`` plus the CF grid-mapping attributes of the CRS.  The
reference writes NetCDF-4 (= HDF5) through xarray; neither xarray nor an HDF5 library exists offline, so this module

* reads the reference's **NetCDF-4 / HDF5** files with the dependency-free reader of :mod:`raider_b200.hdf5_lite`,
* reads/writes the same variable layout as **NetCDF-3 classic** through ``scipy.io.netcdf_file`` (always available),
* reads ``.npz`` archives with the same keys.

The CRS is taken from the CF grid-mapping attributes xarray / pyproj write next to ``crs_wkt`` (``grid_mapping_name`` =
``latitude_longitude`` or ``lambert_conformal_conic`` with its parallels / origin / sphere radius), from a ``proj4`` attribute
(what :func:`write_cube` writes), or -- when pyproj is importable -- from the WKT.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

FIELDS = ('wet', 'hydro', 'wet_total', 'hydro_total')


def _scalar(v):
    return float(np.asarray(v).ravel()[0])


def _crs_from_attrs(attrs: dict):
    from .crs import parse_crs
    if attrs.get('proj4'):
        v = attrs['proj4']
        return parse_crs(v.decode() if isinstance(v, bytes) else v)
    gm = attrs.get('grid_mapping_name')
    gm = gm.decode() if isinstance(gm, bytes) else gm
    if gm == 'latitude_longitude':
        return parse_crs(4326)
    if gm == 'lambert_conformal_conic':
        a = _scalar(attrs.get('semi_major_axis', attrs.get('earth_radius', 6378137.0)))
        b = _scalar(attrs.get('semi_minor_axis', a))
        sp = np.atleast_1d(np.asarray(attrs['standard_parallel'], dtype=np.float64))
        d = dict(proj='lcc', lat_1=float(sp[0]), lat_2=float(sp[-1]), lat_0=_scalar(attrs['latitude_of_projection_origin']),
                 lon_0=_scalar(attrs['longitude_of_central_meridian']), a=a, b=b, x_0=_scalar(attrs.get('false_easting', 0.0)),
                 y_0=_scalar(attrs.get('false_northing', 0.0)))
        return parse_crs(d)
    if gm is not None:
        raise NotImplementedError(f'weather-model grid mapping {gm!r} is not supported by the B200 delay path')
    wkt = attrs.get('crs_wkt')
    if wkt is None:
        return parse_crs(4326)  # delay.py:69-73: warn + assume WGS84
    wkt = wkt.decode() if isinstance(wkt, bytes) else wkt
    try:
        import pyproj
        return parse_crs(pyproj.CRS.from_wkt(wkt))
    except ImportError:
        if 'Lambert' in wkt or 'PROJCRS' in wkt or 'PROJCS' in wkt:
            raise NotImplementedError('projected weather-model CRS given as WKT only needs pyproj to be parsed')
        return parse_crs(4326)


def load_cube(path_or_ds) -> dict:
    """Return {x, y, z, wet, hydro, wet_total, hydro_total, crs} from a file path, an xarray Dataset or a dict."""
    if isinstance(path_or_ds, dict):
        out = dict(path_or_ds)
        out.setdefault('crs', None)
        return out
    if hasattr(path_or_ds, 'variables') and not isinstance(path_or_ds, (str, Path)):  # xarray.Dataset duck type
        ds = path_or_ds
        out = {k: np.array(ds.variables[k][:]) for k in ('x', 'y', 'z')}
        for k in FIELDS:
            if k in ds.variables:
                out[k] = np.array(ds.variables[k][:])
        attrs = dict(getattr(ds['proj'], 'attrs', {})) if 'proj' in ds.variables else {}
        out['crs'] = _crs_from_attrs(attrs)
        return out
    path = Path(path_or_ds)
    if path.suffix == '.npz':
        with np.load(path, allow_pickle=False) as z:
            out = {k: z[k] for k in z.files if k != 'proj4'}
            out['crs'] = _crs_from_attrs({'proj4': str(z['proj4'])} if 'proj4' in z.files else {})
        return out
    with open(path, 'rb') as f:
        magic = f.read(4)
    if magic[:3] == b'CDF':
        from scipy.io import netcdf_file
        with netcdf_file(str(path), 'r', mmap=False) as nc:
            def native(a):   # NetCDF-3 is big-endian on disk
                a = np.array(a)
                return a.astype(a.dtype.newbyteorder('='))
            out = {k: native(nc.variables[k][:]) for k in ('x', 'y', 'z')}
            for k in FIELDS:
                if k in nc.variables:
                    out[k] = native(nc.variables[k][:])
            attrs = dict(nc.variables['proj']._attributes) if 'proj' in nc.variables else {}
        out['crs'] = _crs_from_attrs(attrs)
        return out
    if magic == b'\x89HDF':
        # the reference's own format (weatherModel.py:659-724): NetCDF-4 = HDF5, read without any HDF5 library
        from . import hdf5_lite
        with hdf5_lite.File(path) as f:
            out = {k: np.asarray(f[k].read(), dtype=np.float64) for k in ('x', 'y', 'z')}
            for k in FIELDS:
                if k in f:
                    out[k] = f[k].read()
            attrs = dict(f['proj'].attrs) if 'proj' in f else {}
        out['crs'] = _crs_from_attrs(attrs)
        return out
    raise ValueError(f'{path}: not a NetCDF-3, NetCDF-4 / HDF5 or .npz weather-model cube')


def write_cube(path, cube: dict, proj4: str = '+proj=longlat +datum=WGS84 +no_defs') -> Path:
    """Write a cube as NetCDF-3 classic (or .npz) with the reference's variable layout."""
    path = Path(path)
    if path.suffix == '.npz':
        np.savez(path, proj4=np.array(proj4), **{k: v for k, v in cube.items() if k in FIELDS + ('x', 'y', 'z')})
        return path
    from scipy.io import netcdf_file
    with netcdf_file(str(path), 'w', version=2) as nc:
        for d in ('z', 'y', 'x'):
            nc.createDimension(d, int(np.size(cube[d])))
            v = nc.createVariable(d, 'f8', (d,))
            v[:] = np.asarray(cube[d], dtype=np.float64)
        for k in FIELDS:
            if k in cube:
                # wet / hydro float32, the zenith totals float64 -- the dtypes of the reference's files (weatherModel.py:398-403,617-619)
                dt = ('f8', np.float64) if k.endswith('_total') else ('f4', np.float32)
                v = nc.createVariable(k, dt[0], ('z', 'y', 'x'))
                v[:] = np.asarray(cube[k], dtype=dt[1])
        p = nc.createVariable('proj', 'i4', ())
        p.data[()] = 0  # (netcdf_variable.assignValue indexes [:], which a 0-d array rejects)
        p.proj4 = proj4
        p.crs_wkt = 'GEOGCRS["WGS 84"]' if 'longlat' in proj4 else 'PROJCRS["custom"]'
    return path
