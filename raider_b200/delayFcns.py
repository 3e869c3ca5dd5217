"""Interpolator factory of the delay path (reference: tools/RAiDER/delayFcns.py:23-58).

``getInterpolators`` keeps its name, arguments and return shape -- a pair ``(ifWet, ifHydro)`` whose objects expose
``.grid == (ys, xs, zs)`` (read at delay.py:239) and ``__call__(pts[..., 3]) -> [...]`` with scipy's
``fill_value=nan, bounds_error=False`` linear semantics -- but both objects are thin views on ONE cube staged in HBM
(:class:`raider_b200.engine.DeviceCube`), so a ray-tracing call samples wet and hydrostatic refractivity in one pass.
"""
from __future__ import annotations

import logging
from pathlib import Path
from typing import Union

import numpy as np

from . import _lib
from .cube_io import load_cube
from .engine import DeviceCube

logger = logging.getLogger('RAiDER')


class DeviceInterpolator:
    """One field (0 = wet, 1 = hydro) of a :class:`DeviceCube`, callable like scipy's RegularGridInterpolator."""

    def __init__(self, cube: DeviceCube, field: int) -> None:
        self.cube = cube
        self.field = field
        self.fill_value = np.nan
        self.bounds_error = False
        self.method = 'linear'

    @property
    def grid(self):
        return self.cube.grid

    def __call__(self, xi):
        return self.cube.sample(xi)[self.field]


def getInterpolators(wm_file: Union[dict, Path, str, DeviceCube], kind: str = 'pointwise', shared: bool = False, device=None):
    """Stage the cube of a processed weather model file and return (ifWet, ifHydro).

    ``wm_file`` may be a path (NetCDF-4 / HDF5 as the reference writes it, NetCDF-3 classic, .npz), an xarray Dataset, or a
    dict with keys x, y, z, wet, hydro[, wet_total, hydro_total] holding (z, y, x) arrays.  ``kind='total'`` selects
    the ``*_total`` fields (zenith path), ``kind='ztd'`` is the reference's point-mode re-interpolation of a *delay*
    cube (delay.py:116), which also reads ``wet``/``hydro``.  ``shared`` (a multiprocessing stub in the reference,
    delayFcns.py:46-53) is accepted and ignored: the cube lives once in HBM.
    """
    if isinstance(wm_file, DeviceCube):
        cube = wm_file
    else:
        ds = load_cube(wm_file)
        wet = np.asarray(ds['wet_total' if kind == 'total' else 'wet'])
        hydro = np.asarray(ds['hydro_total' if kind == 'total' else 'hydro'])
        if kind == 'pointwise':   # the refractivity fields are float32 at rest (weatherModel.py:617-619)
            wet, hydro = wet.astype(np.float32, copy=False), hydro.astype(np.float32, copy=False)
        if np.any(np.isnan(wet)) or np.any(np.isnan(hydro)):
            logger.critical('Weather model contains NaNs!')
        cube = DeviceCube(ds['y'], ds['x'], ds['z'], wet, hydro, layout=_lib.LAYOUT_ZYX, crs=ds.get('crs'), device=device)
    return DeviceInterpolator(cube, 0), DeviceInterpolator(cube, 1)


def as_device_cube(interpolators, crs=None) -> DeviceCube:
    """The DeviceCube behind a pair of interpolators; scipy RGIs (from the reference's own factory) are staged on the fly."""
    a, b = interpolators[0], interpolators[1]
    if isinstance(a, DeviceInterpolator) and isinstance(b, DeviceInterpolator) and a.cube is b.cube and (a.field, b.field) == (0, 1):
        return a.cube
    if isinstance(a, DeviceInterpolator) or isinstance(b, DeviceInterpolator):
        raise TypeError('interpolators must be the (ifWet, ifHydro) pair returned by one getInterpolators call')
    ys, xs, zs = (np.asarray(g) for g in a.grid)
    # (float64 values are staged as hi + lo float32 parts by DeviceCube: the zenith path stays at scipy's precision)
    return DeviceCube(ys, xs, zs, np.asarray(a.values), np.asarray(b.values), layout=_lib.LAYOUT_YXZ, crs=crs)
