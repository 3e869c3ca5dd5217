"""Geodesy helpers of the delay path (reference: tools/RAiDER/utilFcns.py:67-137).

``lla2ecef`` / ``ecef2lla`` run on the device (the reference goes through PROJ); the ENU rotations are the
reference's own few-line NumPy formulas (they only ever see one constant vector or a LOS raster on the host).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check, f64, ptr


def sind(x):
    """Return the sine of x when x is in degrees (utilFcns.py:67-69)."""
    return np.sin(np.radians(x))


def cosd(x):
    """Return the cosine of x when x is in degrees (utilFcns.py:72-74)."""
    return np.cos(np.radians(x))


def _bcast3(a, b, c):
    a, b, c = np.broadcast_arrays(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), np.asarray(c, dtype=np.float64))
    return f64(a).ravel(), f64(b).ravel(), f64(c).ravel(), a.shape


def lla2ecef(lat, lon, height, device=None):
    """Transforms from lla to ecef (utilFcns.py:77-81) -> (x, y, z)."""
    la, lo, h, shape = _bcast3(lat, lon, height)
    x, y, z = np.empty(la.size), np.empty(la.size), np.empty(la.size)
    check(_lib.load().rdr_lla2ecef(ptr(la), ptr(lo), ptr(h), la.size, ptr(x), ptr(y), ptr(z), _lib.default_device() if device is None else device))
    if shape == ():
        return float(x[0]), float(y[0]), float(z[0])
    return x.reshape(shape), y.reshape(shape), z.reshape(shape)


def ecef2lla(x, y, z, device=None):
    """Converts ecef to lla (utilFcns.py:84-88).  As in the reference (always_xy=True) the tuple is (lon, lat, height)."""
    xx, yy, zz, shape = _bcast3(x, y, z)
    lon, lat, h = np.empty(xx.size), np.empty(xx.size), np.empty(xx.size)
    check(_lib.load().rdr_ecef2lla(ptr(xx), ptr(yy), ptr(zz), xx.size, ptr(lon), ptr(lat), ptr(h), _lib.default_device() if device is None else device))
    if shape == ():
        return float(lon[0]), float(lat[0]), float(h[0])
    return lon.reshape(shape), lat.reshape(shape), h.reshape(shape)


def enu2ecef(east, north, up, lat0, lon0, h0=None):
    """Converts enu to ecef (utilFcns.py:91-121)."""
    t = cosd(lat0) * up - sind(lat0) * north
    w = sind(lat0) * up + cosd(lat0) * north
    u = cosd(lon0) * t - sind(lon0) * east
    v = sind(lon0) * t + cosd(lon0) * east
    return np.stack((u, v, w), axis=-1)


def ecef2enu(xyz, lat, lon, height=None):
    """Convert ECEF xyz to ENU (utilFcns.py:124-137)."""
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    t = cosd(lon) * x + sind(lon) * y
    e = -sind(lon) * x + cosd(lon) * y
    n = -sind(lat) * t + cosd(lat) * z
    u = cosd(lat) * t + sind(lat) * z
    return np.stack((e, n, u), axis=-1)
