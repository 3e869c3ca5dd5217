"""Reading/writing processed weather-model cubes (the file the delay path consumes).

On-disk contract (reference writer tools/RAiDER/models/weatherModel.py:659-724; readers tools/RAiDER/delay.py:66-78,
tools/RAiDER/delayFcns.py:31-41): variables ``wet, hydro, wet_total, hydro_total`` with dims (z, y, x), coordinate
variables ``x, y, z`` and a ``proj`` variable carrying ``crs_wkt``.  The reference writes NetCDF-4/HDF5 through
xarray; neither xarray nor an HDF5 reader exists offline, so this module

* reads/writes the same variable layout as **NetCDF-3 classic** through ``scipy.io.netcdf_file`` (always available),
* reads ``.npz`` archives with the same keys,
* and, when ``xarray`` *is* importable (a real RAiDER environment), opens NetCDF-4 files through it.

The CRS travels as a proj4 string attribute ``proj4`` on ``proj`` next to ``crs_wkt`` when we write; when reading a
file produced by the reference (WKT only) the WKT is handed to pyproj if present, else EPSG:4326 is assumed unless the
WKT names a Lambert conic (then pyproj is required and we say so).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

FIELDS = ('wet', 'hydro', 'wet_total', 'hydro_total')


def _crs_from_attrs(attrs: dict):
    from .crs import parse_crs
    if attrs.get('proj4'):
        v = attrs['proj4']
        return parse_crs(v.decode() if isinstance(v, bytes) else v)
    wkt = attrs.get('crs_wkt')
    if wkt is None:
        return parse_crs(4326)  # delay.py:69-73: warn + assume WGS84
    wkt = wkt.decode() if isinstance(wkt, bytes) else wkt
    try:
        import pyproj
        return parse_crs(pyproj.CRS.from_wkt(wkt))
    except ImportError:
        if 'Lambert' in wkt or 'PROJCRS' in wkt or 'PROJCS' in wkt:
            raise NotImplementedError('projected weather-model CRS given as WKT needs pyproj to be parsed')
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
            out = {k: np.array(nc.variables[k][:]) for k in ('x', 'y', 'z')}
            for k in FIELDS:
                if k in nc.variables:
                    out[k] = np.array(nc.variables[k][:])
            attrs = dict(nc.variables['proj']._attributes) if 'proj' in nc.variables else {}
        out['crs'] = _crs_from_attrs(attrs)
        return out
    try:
        import xarray as xr
    except ImportError as e:
        raise ImportError(f'{path} is NetCDF-4/HDF5; reading it needs xarray (+h5netcdf/netCDF4), which is not installed. '
                          'Convert it to NetCDF-3 classic or .npz, or install xarray.') from e
    with xr.load_dataset(path) as ds:
        return load_cube(ds)


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
                v = nc.createVariable(k, 'f4', ('z', 'y', 'x'))
                v[:] = np.asarray(cube[k], dtype=np.float32)
        p = nc.createVariable('proj', 'i4', ())
        p.data[()] = 0  # (netcdf_variable.assignValue indexes [:], which a 0-d array rejects)
        p.proj4 = proj4
        p.crs_wkt = 'GEOGCRS["WGS 84"]' if 'longlat' in proj4 else 'PROJCRS["custom"]'
    return path
