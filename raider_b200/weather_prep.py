"""Weather-model processing on the device: the step that *produces* the cube the delay path reads.

Reference: ``WeatherModel.load`` after ``load_weather`` (tools/RAiDER/models/weatherModel.py:252-260) -- ``_find_e`` (:333-354,
``find_svp`` :750-780), ``_uniform_in_z`` (:603-623, three ``interpolate_along_axis`` calls), ``_checkForNans`` (:625-629),
``_get_wet_refractivity`` / ``_get_hydro_refractivity`` (:355-361), ``_adjust_grid`` (:371-387) and ``_getZTD`` (:389-403).
All of it runs in one kernel (K7, a warp per model column); the Python here only shapes arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, f64, ptr
from .constants import _ZMIN


def process_weather(zs, p, t, hum, zlevels=None, k1=0.776, k2=0.233, k3=3.75e3, humidity_type='q', zmin=_ZMIN, xs=None, ys=None,
                    keep_pte=False, device=None) -> dict:
    """Native model fields -> processed cube variables.

    ``zs, p, t, hum``: (ny, nx, nl) arrays as the reference holds them after ``load_weather`` (heights per column, pressure
    [Pa], temperature [K], specific humidity or relative humidity [%]); ``zlevels``: target heights (default: the mean
    column, weatherModel.py:608-612).  Returns ``{z, wet, hydro, wet_total, hydro_total[, p, t, e][, x, y]}`` with (z, y, x)
    float32 fields -- the layout of the processed weather-model file (weatherModel.py:676-724) that ``getInterpolators`` reads.
    """
    if humidity_type not in ('q', 'rh'):
        raise RuntimeError('Not a valid humidity type')
    zs, p, t, hum = (np.asarray(a, dtype=np.float64) for a in (zs, p, t, hum))
    if zs.ndim != 3 or not (zs.shape == p.shape == t.shape == hum.shape):
        raise TypeError("'zs', 'p', 't' and 'hum' must be (ny, nx, nz) arrays of one shape")
    ny, nx, nl = zs.shape
    if zlevels is None:
        zlevels = np.nanmean(zs, axis=(0, 1))
    zlevels = f64(zlevels)
    if np.any(np.diff(zs, axis=2) <= 0):  # interpolate_along_axis sorts unsorted columns (assume_sorted=False); do it once here
        order = np.argsort(zs, axis=2, kind='stable')
        zs, p, t, hum = (np.take_along_axis(a, order, axis=2) for a in (zs, p, t, hum))
    zs, p, t, hum = (f64(a) for a in (zs, p, t, hum))
    pad = 1 if zmin < zlevels[0] else 0
    nzo = zlevels.size + pad
    outs = [np.empty((ny, nx, nzo), dtype=np.float32) for _ in range(7 if keep_pte else 4)]
    nz_written = C.c_int64(0)
    extra = [ptr(o) for o in outs[4:]] if keep_pte else [None, None, None]
    check(_lib.load().rdr_prepare_cube(ny * nx, nl, ptr(zs), ptr(p), ptr(t), ptr(hum), int(humidity_type == 'rh'), ptr(zlevels), zlevels.size,
                                       float(k1), float(k2), float(k3), float(zmin), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(outs[3]),
                                       *extra, C.byref(nz_written), _lib.default_device() if device is None else device, _lib.MEM_HOST))
    assert nz_written.value == nzo
    tr = lambda a: np.ascontiguousarray(np.moveaxis(a, 2, 0))
    cube = {'z': np.concatenate([[float(zmin)], zlevels]) if pad else zlevels.copy(), 'wet': tr(outs[0]), 'hydro': tr(outs[1]),
            'wet_total': tr(outs[2]), 'hydro_total': tr(outs[3])}
    if keep_pte:
        cube.update(p=tr(outs[4]), t=tr(outs[5]), e=tr(outs[6]))
    if xs is not None:
        cube['x'] = np.unique(np.asarray(xs, dtype=np.float64))  # weatherModel.py:621-622
    if ys is not None:
        cube['y'] = np.unique(np.asarray(ys, dtype=np.float64))
    return cube
