"""WGS-84 geodesy for the oracle -- TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference does every geodetic<->geocentric conversion through PROJ:
``pyproj.Transformer.from_crs(4326, 4978, always_xy=True)`` and its inverse
(tools/RAiDER/utilFcns.py:77-88, tools/RAiDER/delay.py:238,253,267,295).  PROJ is a
third-party dependency that is NOT vendored in /root/reference and is not installed
offline (``pyproj>=2.2.0`` in environment.yml:34, PROJ version unpinned), so its published
algorithm is restated here: the ``cart`` conversion (PROJ ``src/conversions/cart.cpp``),
which is

* forward  (geodetic -> cartesian):  N = a / sqrt(1 - es sin^2(phi));
  x = (N+h) cos(phi) cos(lam), y = (N+h) cos(phi) sin(lam), z = (N (1-es) + h) sin(phi)
* inverse  (cartesian -> geodetic):  Bowring's (1976) single-step closed form with the
  normalised parametric-latitude terms PROJ uses (no iteration), height = p / cos(phi) - N,
  and the polar branch ``|z| - geocentric_radius`` when cos(phi) < 1e-6,

followed by the ``unitconvert`` rad<->deg step (x RAD_TO_DEG = 57.29577951308232).

Parity with PROJ itself is UNPINNED below ~1e-9 m (the reference only pins this boundary with
``np.allclose`` at a handful of points, test/test_delayFcns.py:48-99).  What *is* checked in
tests/: round trips, the equator/pole known answers of test_delayFcns.py:86-99, and agreement
with an iterated (converged) solution to < 1e-8 m for |h| < 100 km.
"""
from __future__ import annotations

import numpy as np

# WGS-84 (EPSG:7030) defining constants
WGS84_A = 6378137.0
WGS84_RF = 298.257223563
WGS84_F = 1.0 / WGS84_RF
WGS84_B = WGS84_A * (1.0 - WGS84_F)
WGS84_ES = 2.0 * WGS84_F - WGS84_F * WGS84_F            # first eccentricity squared
WGS84_E2S = WGS84_ES / (1.0 - WGS84_ES)                  # second eccentricity squared
WGS84_B_DIV_A_SQ = (1.0 - WGS84_F) * (1.0 - WGS84_F)
RAD_TO_DEG = 57.295779513082321
DEG_TO_RAD = 0.017453292519943296


def sind(x):
    """utilFcns.py:67-69"""
    return np.sin(np.radians(x))


def cosd(x):
    """utilFcns.py:72-74"""
    return np.cos(np.radians(x))


def _normal_radius_of_curvature(sinphi):
    return WGS84_A / np.sqrt(1.0 - WGS84_ES * sinphi * sinphi)


def lla2ecef(lat, lon, height):
    """utilFcns.py:77-81: ``T(4326->4978, always_xy).transform(lon, lat, height)`` -> (x, y, z)."""
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    height = np.asarray(height, dtype=np.float64)
    phi = lat * DEG_TO_RAD
    lam = lon * DEG_TO_RAD
    cosphi = np.cos(phi)
    sinphi = np.sin(phi)
    N = _normal_radius_of_curvature(sinphi)
    x = (N + height) * cosphi * np.cos(lam)
    y = (N + height) * cosphi * np.sin(lam)
    z = (N * (1.0 - WGS84_ES) + height) * sinphi
    return x, y, z


def _bowring(x, y, z):
    """Shared core of the inverse: returns (p, x_phi, y_phi, cosphi, sinphi)."""
    p = np.hypot(x, y)
    y_theta = z * WGS84_A
    x_theta = p * WGS84_B
    norm = np.hypot(y_theta, x_theta)
    with np.errstate(invalid='ignore', divide='ignore'):
        c = np.where(norm == 0, 1.0, x_theta / norm)
        s = np.where(norm == 0, 0.0, y_theta / norm)
        y_phi = z + WGS84_E2S * WGS84_B * s * s * s
        x_phi = p - WGS84_ES * WGS84_A * c * c * c
        norm_phi = np.hypot(y_phi, x_phi)
        cosphi = np.where(norm_phi == 0, 1.0, x_phi / norm_phi)
        sinphi = np.where(norm_phi == 0, 0.0, y_phi / norm_phi)
    return p, x_phi, y_phi, cosphi, sinphi


def _height(p, z, cosphi, sinphi):
    with np.errstate(invalid='ignore', divide='ignore'):
        h_reg = p / cosphi - _normal_radius_of_curvature(sinphi)
        c2 = cosphi * cosphi
        s2 = sinphi * sinphi
        bs2 = WGS84_B_DIV_A_SQ * s2
        r = WGS84_A * np.sqrt((c2 + WGS84_B_DIV_A_SQ * bs2) / (c2 + bs2))
        h_pole = np.abs(z) - r
    return np.where(cosphi < 1e-6, h_pole, h_reg)


def ecef2height(x, y, z):
    """Height component only of :func:`ecef2lla` (what getTopOfAtmosphere consumes, losreader.py:730-731)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    p, x_phi, y_phi, cosphi, sinphi = _bowring(x, y, z)
    pole = x_phi <= 0
    cosphi = np.where(pole, 0.0, cosphi)
    sinphi = np.where(pole, np.where(z >= 0, 1.0, -1.0), sinphi)
    return _height(p, z, cosphi, sinphi)


def ecef2lla(x, y, z):
    """utilFcns.py:84-88: ``T(4978->4326, always_xy).transform(x, y, z)`` -> (lon_deg, lat_deg, h).

    NB the name says "lla" but with ``always_xy=True`` PROJ returns longitude first; the
    reference only ever reads element [2] of the result (losreader.py:731).
    """
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    p, x_phi, y_phi, cosphi, sinphi = _bowring(x, y, z)
    pole = x_phi <= 0
    with np.errstate(invalid='ignore', divide='ignore'):
        phi = np.where(pole, np.where(z >= 0, 0.5 * np.pi, -0.5 * np.pi), np.arctan(y_phi / x_phi))
    cosphi = np.where(pole, 0.0, cosphi)
    sinphi = np.where(pole, np.where(z >= 0, 1.0, -1.0), sinphi)
    lam = np.arctan2(y, x)
    h = _height(p, z, cosphi, sinphi)
    return lam * RAD_TO_DEG, phi * RAD_TO_DEG, h


def enu2ecef(east, north, up, lat0, lon0, h0=None):
    """utilFcns.py:91-121 (rotation only; h0 is unused there too)."""
    t = cosd(lat0) * up - sind(lat0) * north
    w = sind(lat0) * up + cosd(lat0) * north
    u = cosd(lon0) * t - sind(lon0) * east
    v = sind(lon0) * t + cosd(lon0) * east
    return np.stack((u, v, w), axis=-1)


def ecef2enu(xyz, lat, lon, height=None):
    """utilFcns.py:124-137."""
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    t = cosd(lon) * x + sind(lon) * y
    e = -sind(lon) * x + cosd(lon) * y
    n = -sind(lat) * t + cosd(lat) * z
    u = cosd(lat) * t + sind(lat) * z
    return np.stack((e, n, u), axis=-1)


def inc_hd_to_enu(incidence, heading):
    """losreader.py:374-396."""
    incidence = np.asarray(incidence, dtype=np.float64)
    heading = np.asarray(heading, dtype=np.float64)
    if np.any(incidence < 0):
        raise ValueError('inc_hd_to_enu: Incidence angle cannot be less than 0')
    east = sind(incidence) * cosd(heading + 90)
    north = sind(incidence) * sind(heading + 90)
    up = cosd(incidence)
    return np.stack((east, north, up), axis=-1)


def getZenithLookVecs(lats, lons, heights=None):
    """losreader.py:302-316."""
    x = np.cos(np.radians(lats)) * np.cos(np.radians(lons))
    y = np.cos(np.radians(lats)) * np.sin(np.radians(lons))
    z = np.sin(np.radians(lats))
    return np.stack([x, y, z], axis=-1)


# ---------------------------------------------------------------------------------------------
# Lambert conformal conic on a sphere (HRRR: tools/RAiDER/models/hrrr.py:255-260 -> PROJ "lcc")
# ---------------------------------------------------------------------------------------------
class LambertConformalSphere:
    """PROJ ``+proj=lcc +lat_1 +lat_2 +lat_0 +lon_0 +R`` (spherical branch of PROJ lcc.cpp), restated.

    HRRR uses lat_1 = lat_2 = lat_0 = 38.5, lon_0 = 262.5, a = b = 6371229 (hrrr.py:250-260).
    Only the forward direction (lon,lat -> x,y) is on the hot path (delay.py:295 with an LCC model crs).
    NOTE: with a *spherical* model CRS the reference's ``Transformer.from_crs(4978, model_crs)``
    goes ECEF -(WGS84 cart inverse)-> geodetic lat/lon/h and then projects those angles on the
    sphere (PROJ applies no datum shift between two "unknown"-datum ellipsoids by default); that
    is what is restated: lat/lon from :func:`ecef2lla`, then this projection, h unchanged.
    """

    def __init__(self, lat_1=38.5, lat_2=38.5, lat_0=38.5, lon_0=262.5, R=6371229.0, x_0=0.0, y_0=0.0):
        self.lat_1, self.lat_2, self.lat_0, self.lon_0 = float(lat_1), float(lat_2), float(lat_0), float(lon_0)
        self.R = float(R)
        self.lam0 = float(lon_0) * DEG_TO_RAD
        self.x_0 = float(x_0)
        self.y_0 = float(y_0)
        phi1 = float(lat_1) * DEG_TO_RAD
        phi2 = float(lat_2) * DEG_TO_RAD
        phi0 = float(lat_0) * DEG_TO_RAD
        sinphi = np.sin(phi1)
        cosphi = np.cos(phi1)
        secant = abs(phi1 - phi2) >= 1e-10
        n = sinphi
        if secant:
            n = np.log(cosphi / np.cos(phi2)) / np.log(np.tan(0.25 * np.pi + 0.5 * phi2) / np.tan(0.25 * np.pi + 0.5 * phi1))
        self.n = float(n)
        self.c = float(cosphi * np.tan(0.25 * np.pi + 0.5 * phi1) ** n / n)
        self.rho0 = 0.0 if abs(abs(phi0) - 0.5 * np.pi) < 1e-10 else float(self.c * np.tan(0.25 * np.pi + 0.5 * phi0) ** (-n))

    def params(self):
        """Packed parameters in the order the C-ABI takes them (include/raider_b200.h)."""
        return np.array([self.n, self.c, self.rho0, self.lam0, self.R, self.x_0, self.y_0], dtype=np.float64)

    def forward(self, lon_deg, lat_deg):
        lam = np.asarray(lon_deg, dtype=np.float64) * DEG_TO_RAD - self.lam0
        # PROJ normalises the longitude difference into [-pi, pi] (adjlon)
        lam = np.where(np.abs(lam) > np.pi, lam - 2.0 * np.pi * np.round(lam / (2.0 * np.pi)), lam)
        phi = np.asarray(lat_deg, dtype=np.float64) * DEG_TO_RAD
        rho = self.c * np.power(np.tan(0.25 * np.pi + 0.5 * phi), -self.n)
        lam = lam * self.n
        x = self.R * (rho * np.sin(lam)) + self.x_0
        y = self.R * (self.rho0 - rho * np.cos(lam)) + self.y_0
        return x, y

    def inverse(self, x, y):
        x = (np.asarray(x, dtype=np.float64) - self.x_0) / self.R
        y = self.rho0 - (np.asarray(y, dtype=np.float64) - self.y_0) / self.R
        rho = np.hypot(x, y)
        sgn = np.sign(self.n) if self.n != 0 else 1.0
        rho_s = rho * sgn
        xs, ys = x * sgn, y * sgn
        phi = 2.0 * np.arctan(np.power(self.c / rho_s, 1.0 / self.n)) - 0.5 * np.pi
        lam = np.arctan2(xs, ys) / self.n
        return (lam + self.lam0) * RAD_TO_DEG, phi * RAD_TO_DEG
