"""CPU restatement of the weather-model processing that produces the cube the delay path reads -- TEST INFRASTRUCTURE
(see oracle/__init__.py).

Follows ``WeatherModel.load`` (tools/RAiDER/models/weatherModel.py:235-261), step for step, on in-memory arrays laid out
(y, x, z) like the reference's:

* :func:`find_svp`            weatherModel.py:750-780   saturation vapour pressure (Buck over water, Alduchov-Eskridge over
                              ice, quadratic blend between -23 C and 0 C), returned as float32
* :func:`find_e`              weatherModel.py:333-354   partial pressure of water vapour from specific / relative humidity
* :func:`uniform_in_z`        weatherModel.py:603-623   three ``interpolate_along_axis(..., fill_value=nan)`` calls, cast to
                              float32 (through :func:`oracle.interp.interpolate_along_axis`, itself pinned bit for bit
                              against the compiled reference native)
* :func:`fill_nans`           weatherModel.py:625-629 -> interpolator.py:110-130 (``fillna3D``): leading NaNs take the first
                              valid value, interior NaNs are interpolated linearly *in index*, trailing NaNs get the fill value
* :func:`refractivity`        weatherModel.py:355-361   k2 e / T + k3 e / T^2 and k1 P / T in the arrays' own precision
* :func:`adjust_grid`         weatherModel.py:371-387   one extra level at ``zmin`` copied from the lowest valid value
* :func:`get_ztd`             weatherModel.py:389-403   1e-6 trapz(field[level:], zs[level:])
* :func:`process`             the sequence above -> the processed-file variables (z, y, x)

Pinned by the reference's own known answers: ``test_uniform_in_z_small`` (test/test_weather_model.py:178-211) and the
``MockWeatherModel`` analytic refractivities / ZTDs (:113-133, asserted at :385-401) -- tests/test_oracle_pins.py.
"""
from __future__ import annotations

import numpy as np

from . import interp

R_V, R_D = 461.524, 287.06  # weatherModel.py:75-76
ZMIN = np.float64(-100)     # constants.py:11


def find_svp(t):
    t = np.asarray(t)
    t1, t2 = 273.15, 250.15
    tref = t - t1
    wgt = (t - t2) / (t1 - t2)
    svpw = 6.1121 * np.exp((17.502 * tref) / (240.97 + tref))
    svpi = 6.1121 * np.exp((22.587 * tref) / (273.86 + tref))
    svp = svpi + (svpw - svpi) * wgt ** 2
    svp = np.where(t > t1, svpw, svp)
    svp = np.where(t < t2, svpi, svp)
    return (svp * 100).astype(np.float32)


def find_e(p, t, hum, humidity_type='q'):
    svp = find_svp(t)
    if humidity_type == 'q':
        w = hum / (1 - hum)
        return w * R_V * (p - svp) / R_D
    if humidity_type == 'rh':
        return hum / 100 * svp
    raise RuntimeError('Not a valid humidity type')


def uniform_in_z(zs, p, t, e, zlevels=None):
    if zlevels is None:
        zlevels = np.nanmean(zs, axis=(0, 1))
    zlevels = np.asarray(zlevels, dtype=np.float64)
    new_zs = np.tile(zlevels, zs.shape[:2] + (1,))
    out = [interp.interpolate_along_axis(zs, v, new_zs, axis=2, fill_value=np.nan).astype(np.float32) for v in (t, p, e)]
    return zlevels, out[1], out[0], out[2]  # zs, p, t, e


def fill_nans(a, fill_value=0.0):
    """fillna3D along the last axis, as pandas' ``interpolate(axis=1, limit_direction='backward')`` + fill does it."""
    a = np.array(a, copy=True)
    flat = a.reshape(-1, a.shape[-1])
    for row in flat:
        ok = np.flatnonzero(~np.isnan(row))
        if ok.size == 0:
            row[:] = fill_value
            continue
        idx = np.arange(row.size)
        inner = (idx > ok[0]) & (idx < ok[-1]) & np.isnan(row)
        row[inner] = np.interp(idx[inner], ok, row[ok])
        row[: ok[0]] = row[ok[0]]
        row[ok[-1] + 1:] = fill_value
    return a


def refractivity(p, t, e, k1, k2, k3):
    wet = k2 * e / t + k3 * e / t ** 2
    hydro = k1 * p / t
    return wet, hydro


def adjust_grid(zs, arrays, zmin=ZMIN):
    if zmin < np.nanmin(zs):
        zs = np.insert(zs, 0, zmin)
        out = []
        for a in arrays:
            first = (~np.isnan(a)).argmax(-1)
            low = np.take_along_axis(a, first[..., None], axis=-1)
            out.append(np.concatenate((low, a), axis=2))
        return zs, out
    return zs, list(arrays)


def get_ztd(zs, field):
    total = np.zeros(field.shape)
    trapz = getattr(np, 'trapezoid', None) or np.trapz
    for level in range(field.shape[2]):
        total[..., level] = 1e-6 * trapz(field[..., level:], x=zs[level:], axis=2)
    return total


def process(zs, p, t, hum, zlevels, k1, k2, k3, humidity_type='q', zmin=ZMIN):
    """The whole of WeatherModel.load after load_weather.  Returns the processed-file variables, fields as (z, y, x)."""
    e = find_e(p, t, hum, humidity_type)
    zl, p, t, e = uniform_in_z(zs, p, t, e, zlevels)
    p, t, e = fill_nans(p), fill_nans(t, fill_value=1e16), fill_nans(e)
    wet, hydro = refractivity(p, t, e, k1, k2, k3)
    zl, (p, t, e, wet, hydro) = adjust_grid(zl, (p, t, e, wet, hydro), zmin)
    wet_total, hydro_total = get_ztd(zl, wet), get_ztd(zl, hydro)
    tr = lambda a: np.ascontiguousarray(np.moveaxis(a, 2, 0))
    return {'z': zl, 'wet': tr(wet), 'hydro': tr(hydro), 'wet_total': tr(wet_total), 'hydro_total': tr(hydro_total),
            'p': tr(p), 't': tr(t), 'e': tr(e)}
