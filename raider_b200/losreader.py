"""Line-of-sight geometry layer (reference: tools/RAiDER/losreader.py).

The LOS classes stay the thin Python entry layer they are in the reference (``is_Zenith`` / ``is_Projected`` /
``ray_trace`` / ``setPoints`` / ``setTime`` / ``getLookVectors`` / ``__call__``).  What moves to the GPU:
``getTopOfAtmosphere`` (:706-733) and ``build_ray`` (:772-835) -- as API-compatible functions here and, fused over
whole rasters, inside ``raider_b200.delay._build_cube_ray``.  LOS classes that can be evaluated per pixel on the device
advertise it through ``device_spec()``; any other object with ``getLookVectors(ht, llh, xyz, yy)`` (e.g. the
reference's isce3-backed ``Raytracing``) is called on the host and its (ny, nx, 3) array is uploaded.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC

import numpy as np

from . import _lib
from ._lib import check, f64, ptr
from .constants import _ZREF
from .utilFcns import cosd, enu2ecef, sind


class LOS(ABC):
    """LOS Class definition for handling look vectors (losreader.py:32-72)."""

    def __init__(self) -> None:
        self._lats, self._lons, self._heights = None, None, None
        self._look_vecs = None
        self._ray_trace = False
        self._is_zenith = False
        self._is_projected = False

    def setPoints(self, lats, lons=None, heights=None) -> None:
        """Set the pixel locations."""
        if (lats is None) and (self._lats is None):
            raise RuntimeError("You haven't given any point locations yet")
        if lons is None:
            llh = lats  # assume points are [lats lons heights]
            self._lats = llh[..., 0]
            self._lons = llh[..., 1]
            self._heights = llh[..., 2]
        elif heights is None:
            self._lats = lats
            self._lons = lons
            self._heights = np.zeros((len(lats), 1))
        else:
            self._lats = lats
            self._lons = lons
            self._heights = heights

    def setTime(self, datetime) -> None:
        self._time = datetime

    def is_Zenith(self):
        return self._is_zenith

    def is_Projected(self):
        return self._is_projected

    def ray_trace(self):
        return self._ray_trace


class Zenith(LOS):
    """Class definition for a "Zenith" object (losreader.py:75-91)."""

    def __init__(self) -> None:
        super().__init__()
        self._is_zenith = True

    def setLookVectors(self) -> None:
        if self._lats is None:
            raise ValueError('Target points not set')
        if self._look_vecs is None:
            self._look_vecs = getZenithLookVecs(self._lats, self._lons, self._heights)

    def __call__(self, delays):
        return delays


class Conventional(LOS):
    """Zenith delay projected with the standard cos(inc) scaling (losreader.py:94-133).

    The reference reads an ISCE LOS raster through rasterio or an orbit file through isce3 (neither available
    offline); here the incidence (deg) is given directly -- a scalar, an array matching the query points, or a
    two-band (incidence, heading) array like the raster ``rio_open`` returns.
    """

    def __init__(self, filename=None, los_convention='isce', time=None, pad=600, incidence=None, heading=None) -> None:
        super().__init__()
        self._file = filename
        self._time = time
        self._pad = pad
        self._is_projected = True
        self._convention = los_convention
        if self._convention.lower() != 'isce':
            raise NotImplementedError()
        self._incidence = incidence
        self._heading = 0.0 if heading is None else heading

    def __call__(self, delays):
        if self._lats is None:
            raise ValueError('Target points not set')
        if self._incidence is None:
            if self._file is None:
                raise ValueError('LOS file not set')
            raise NotImplementedError('reading LOS rasters / orbit files needs rasterio or isce3; pass incidence= instead')
        LOS_enu = inc_hd_to_enu(np.asarray(self._incidence, dtype=np.float64), np.asarray(self._heading, dtype=np.float64))
        delays = np.asarray(delays)
        if delays.shape == LOS_enu.shape:
            return delays / LOS_enu
        else:
            return delays / LOS_enu[..., -1]


class Raytracing(LOS):
    """Full ray tracing (losreader.py:136-299) with look vectors the device can generate or that are given explicitly.

    * ``Raytracing(incidence=30, heading=-168)``: constant incidence/heading -> ENU (inc_hd_to_enu) -> ECEF per
      pixel (enu2ecef), evaluated inside the kernels (no (ny,nx,3) array ever exists);
    * ``Raytracing(look_vecs=array)``: explicit ECEF unit vectors, shape (ny, nx, 3) per height or a callable
      ``f(ht, llh, xyz, yy)``;
    * orbit files need isce3 in the reference (``geo2rdr`` per pixel, :219-255) -- not available offline.
    """

    def __init__(self, filename=None, los_convention='isce', time=None, look_dir='right', pad=600, incidence=None, heading=None,
                 look_vecs=None) -> None:
        super().__init__()
        self._ray_trace = True
        self._file = filename
        self._time = time
        self._pad = pad
        self._convention = los_convention
        if self._convention.lower() != 'isce':
            raise NotImplementedError()
        if look_dir.lower() not in ('right', 'left'):
            raise RuntimeError(f'Unknown look direction: {look_dir}')
        self._incidence, self._heading, self._vecs = incidence, heading, look_vecs
        if incidence is None and look_vecs is None:
            raise NotImplementedError('orbit-file look vectors need isce3 (geo2rdr); pass incidence=/heading= or look_vecs=')
        if incidence is not None:
            self._enu = inc_hd_to_enu(np.float64(incidence), np.float64(0.0 if heading is None else heading))

    def device_spec(self):
        if self._incidence is not None and np.ndim(self._incidence) == 0:
            return _lib.LOS_ENU_CONST, f64(self._enu)
        return None

    def getLookVectors(self, ht, llh, xyz, yy):
        """(ny, nx, 3) ECEF unit vectors ground -> sensor (contract of losreader.py:219-255)."""
        if self._vecs is not None:
            v = self._vecs(ht, llh, xyz, yy) if callable(self._vecs) else self._vecs
            return np.asarray(v, dtype=np.float64)
        e, n, u = self._enu
        return enu2ecef(e, n, u, llh[1], llh[0], llh[2])


class ZenithRaytracing(Raytracing):
    """Ray tracing straight up (local zenith look vectors, losreader.py:302-316): integrates refractivity over height."""

    def __init__(self) -> None:
        LOS.__init__(self)
        self._ray_trace = True
        self._incidence = self._heading = self._vecs = None

    def device_spec(self):
        return _lib.LOS_ZENITH, None

    def getLookVectors(self, ht, llh, xyz, yy):
        return getZenithLookVecs(llh[1], llh[0], llh[2])


def getZenithLookVecs(lats, lons, heights):
    """Look vectors when Zenith is used (losreader.py:302-316): (in_shape) x 3 ECEF unit vectors."""
    x = np.cos(np.radians(lats)) * np.cos(np.radians(lons))
    y = np.cos(np.radians(lats)) * np.sin(np.radians(lons))
    z = np.sin(np.radians(lats))
    return np.stack([x, y, z], axis=-1)


def inc_hd_to_enu(incidence, heading):
    """Incidence/heading (deg) -> local ENU unit vector ground -> sensor (losreader.py:374-396)."""
    if np.any(incidence < 0):
        raise ValueError('inc_hd_to_enu: Incidence angle cannot be less than 0')
    east = sind(incidence) * cosd(heading + 90)
    north = sind(incidence) * sind(heading + 90)
    up = cosd(incidence)
    return np.stack((east, north, up), axis=-1)


def getTopOfAtmosphere(xyz, look_vecs, toaheight, factor=None, device=None):
    """Ray intersection with geodetic height ``toaheight`` by Newton-Raphson (losreader.py:706-733), on the device."""
    xyz = np.asarray(xyz, dtype=np.float64)
    look_vecs = np.asarray(look_vecs, dtype=np.float64)
    shape = np.broadcast_shapes(xyz.shape, look_vecs.shape)
    x = f64(np.broadcast_to(xyz, shape)).reshape(-1, 3)
    u = f64(np.broadcast_to(look_vecs, shape)).reshape(-1, 3)
    fac = None
    if factor is not None:
        fac = f64(np.broadcast_to(np.asarray(factor, dtype=np.float64), shape[:-1])).ravel()
    out = np.empty_like(x)
    check(_lib.load().rdr_top_of_atmosphere(ptr(x), ptr(u), x.shape[0], float(toaheight), ptr(fac), ptr(out),
                                            _lib.default_device() if device is None else device))
    return out.reshape(shape)


def build_ray(model_zs, ht, xyz, LOS, MAX_TROPO_HEIGHT=_ZREF, device=None):
    """Ray length in ECEF between weather-model layers (losreader.py:772-835), on the device.

    Returns ``(ray_lengths (K, ...), low_xyzs (K, ..., 3), high_xyzs (K, ..., 3))`` or ``(None, None, None)`` when no
    layer contributes (:832-833).
    """
    model_zs = f64(model_zs)
    xyz = np.asarray(xyz, dtype=np.float64)
    LOS = np.asarray(LOS, dtype=np.float64)
    shape = np.broadcast_shapes(xyz.shape, LOS.shape)
    x = f64(np.broadcast_to(xyz, shape)).reshape(-1, 3)
    u = f64(np.broadcast_to(LOS, shape)).reshape(-1, 3)
    n = x.shape[0]
    lib = _lib.load()
    K = C.c_int64(0)
    dev = _lib.default_device() if device is None else device
    rc = lib.rdr_build_ray(ptr(model_zs), model_zs.size, float(ht), None, None, 0, float(MAX_TROPO_HEIGHT), C.byref(K), None, None, None, dev)
    if rc == _lib.RDR_ERR_NO_LAYERS:
        return None, None, None
    check(rc)
    k = K.value
    lens, lows, highs = np.empty((k, n)), np.empty((k, n, 3)), np.empty((k, n, 3))
    check(lib.rdr_build_ray(ptr(model_zs), model_zs.size, float(ht), ptr(x), ptr(u), n, float(MAX_TROPO_HEIGHT), C.byref(K), ptr(lens),
                            ptr(lows), ptr(highs), dev))
    return lens.reshape((k,) + shape[:-1]), lows.reshape((k,) + shape), highs.reshape((k,) + shape)
