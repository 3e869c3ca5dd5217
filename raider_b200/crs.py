"""Minimal CRS layer for the delay path.

The reference hands pyproj ``CRS`` objects around (tools/RAiDER/delay.py:63-73,238,251-253).  pyproj/PROJ is not a
dependency of this package (it is not even installable offline), and the hot path needs exactly two families:
geographic WGS-84 (EPSG:4326: ERA5, GMAO, HRES, MERRA2, ...) and HRRR's spherical Lambert conformal conic
(tools/RAiDER/models/hrrr.py:248-260).  ``parse_crs`` accepts what the reference's callers pass -- EPSG ints,
'EPSG:4326' strings, proj4 strings/dicts, pyproj CRS objects (duck-typed via ``to_epsg``/``to_dict``) -- and maps
them onto those two families; anything else raises ``NotImplementedError`` rather than silently mis-projecting.
Host-side forward/inverse here only serve the *query-grid* conversion of delay.py:262-265 (geometry entry layer);
the per-sample ECEF -> model transform of delay.py:295 runs on the device.
"""
from __future__ import annotations

import numpy as np

from . import _lib

_DEG = np.pi / 180.0


class Geographic:
    """EPSG:4326 (x = lon deg, y = lat deg)."""
    kind = _lib.CRS_GEOGRAPHIC

    def params(self):
        return None

    def to_llh(self, xx, yy, hh):
        return [xx, yy, hh]

    def from_ll(self, lon, lat):
        return lon, lat

    def __eq__(self, other):
        try:
            return isinstance(parse_crs(other), Geographic)
        except (NotImplementedError, TypeError):
            return False

    def __repr__(self):
        return 'Geographic(EPSG:4326)'


class LambertConformalSphere:
    """``+proj=lcc +lat_1 +lat_2 +lat_0 +lon_0 +a=R +b=R`` (spherical branch of PROJ's lcc)."""
    kind = _lib.CRS_LCC_SPHERE

    def __init__(self, lat_1=38.5, lat_2=38.5, lat_0=38.5, lon_0=262.5, R=6371229.0, x_0=0.0, y_0=0.0) -> None:
        self.args = dict(lat_1=float(lat_1), lat_2=float(lat_2), lat_0=float(lat_0), lon_0=float(lon_0), R=float(R),
                         x_0=float(x_0), y_0=float(y_0))
        phi1, phi2, phi0 = float(lat_1) * _DEG, float(lat_2) * _DEG, float(lat_0) * _DEG
        n = np.sin(phi1)
        if abs(phi1 - phi2) >= 1e-10:
            n = np.log(np.cos(phi1) / np.cos(phi2)) / np.log(np.tan(0.25 * np.pi + 0.5 * phi2) / np.tan(0.25 * np.pi + 0.5 * phi1))
        self.n = float(n)
        self.c = float(np.cos(phi1) * np.tan(0.25 * np.pi + 0.5 * phi1) ** n / n)
        self.rho0 = 0.0 if abs(abs(phi0) - 0.5 * np.pi) < 1e-10 else float(self.c * np.tan(0.25 * np.pi + 0.5 * phi0) ** (-n))
        self.lam0 = float(lon_0) * _DEG
        self.R, self.x_0, self.y_0 = float(R), float(x_0), float(y_0)

    def params(self):
        return np.array([self.n, self.c, self.rho0, self.lam0, self.R, self.x_0, self.y_0], dtype=np.float64)

    def from_ll(self, lon, lat):
        lam = np.asarray(lon, dtype=np.float64) * _DEG - self.lam0
        lam = np.where(np.abs(lam) > np.pi, lam - 2.0 * np.pi * np.round(lam / (2.0 * np.pi)), lam)
        rho = self.c * np.power(np.tan(0.25 * np.pi + 0.5 * np.asarray(lat, dtype=np.float64) * _DEG), -self.n)
        lam = lam * self.n
        return self.R * (rho * np.sin(lam)) + self.x_0, self.R * (self.rho0 - rho * np.cos(lam)) + self.y_0

    def to_llh(self, xx, yy, hh):
        x = (np.asarray(xx, dtype=np.float64) - self.x_0) / self.R
        y = self.rho0 - (np.asarray(yy, dtype=np.float64) - self.y_0) / self.R
        sgn = -1.0 if self.n < 0 else 1.0
        rho = np.hypot(x, y) * sgn
        phi = 2.0 * np.arctan(np.power(self.c / rho, 1.0 / self.n)) - 0.5 * np.pi
        lam = np.arctan2(x * sgn, y * sgn) / self.n
        return [(lam + self.lam0) / _DEG, phi / _DEG, hh]

    def __eq__(self, other):
        try:
            o = parse_crs(other)
        except (NotImplementedError, TypeError):
            return False
        return isinstance(o, LambertConformalSphere) and o.args == self.args

    def __repr__(self):
        return f'LambertConformalSphere({self.args})'


def _from_proj_dict(d: dict):
    proj = d.get('proj')
    if proj in ('longlat', 'latlong', 'lonlat'):
        # only WGS-84 is geographic here: another datum / ellipsoid would need a datum shift the device geodesy does not do
        datum, ellps = str(d.get('datum', 'WGS84')).upper(), str(d.get('ellps', 'WGS84')).upper()
        if datum != 'WGS84' or ellps != 'WGS84' or 'towgs84' in d or 'nadgrids' in d:
            raise NotImplementedError(f'geographic CRS on datum {d.get("datum", d.get("ellps"))!r}: only WGS-84 (EPSG:4326) is supported')
        if 'a' in d or 'b' in d or 'rf' in d or 'R' in d:
            a, rf = float(d.get('a', d.get('R', 6378137.0))), float(d.get('rf', 298.257223563 if 'R' not in d else 0.0))
            if abs(a - 6378137.0) > 1e-6 or abs(rf - 298.257223563) > 1e-9:
                raise NotImplementedError('geographic CRS on a non-WGS-84 ellipsoid is not supported')
        return Geographic()
    if proj == 'lcc':
        if not any(k in d for k in ('a', 'R')):
            # PROJ's default ellipsoid is GRS80: an lcc without an explicit sphere is ellipsoidal
            raise NotImplementedError('Lambert conformal conic without an explicit sphere (+R or +a == +b): PROJ would use the GRS80 ellipsoid, '
                                      'only the spherical HRRR form is supported')
        a = float(d.get('a', d.get('R')))
        b = float(d.get('b', a))
        if 'R' not in d and abs(a - b) > 1e-6 * a and 'rf' not in d:
            raise NotImplementedError('ellipsoidal Lambert conformal conic is not supported (only the spherical HRRR form)')
        if 'rf' in d or 'ellps' in d or 'datum' in d:
            raise NotImplementedError('ellipsoidal Lambert conformal conic is not supported (only the spherical HRRR form)')
        lat_1 = float(d.get('lat_1', d.get('lat_0', 0.0)))
        return LambertConformalSphere(lat_1=lat_1, lat_2=float(d.get('lat_2', lat_1)), lat_0=float(d.get('lat_0', 0.0)),
                                      lon_0=float(d.get('lon_0', 0.0)), R=a, x_0=float(d.get('x_0', 0.0)), y_0=float(d.get('y_0', 0.0)))
    raise NotImplementedError(f'unsupported projection for the B200 delay path: {d}')


def parse_crs(crs):
    """Accepts Geographic/LambertConformalSphere, EPSG int or 'EPSG:4326', proj4 string/dict, or a pyproj-like CRS."""
    if isinstance(crs, (Geographic, LambertConformalSphere)):
        return crs
    if crs is None:
        return Geographic()
    if isinstance(crs, (int, np.integer)):
        if int(crs) == 4326:
            return Geographic()
        raise NotImplementedError(f'EPSG:{int(crs)} is not supported by the B200 delay path (4326 and spherical LCC only)')
    if isinstance(crs, str):
        s = crs.strip()
        if s.upper().startswith('EPSG:'):
            return parse_crs(int(s.split(':')[-1]))
        if s.isdigit():
            return parse_crs(int(s))
        if '+proj' in s:
            d = {}
            for tok in s.split():
                if tok.startswith('+') and '=' in tok:
                    k, v = tok[1:].split('=', 1)
                    d[k] = v
            return _from_proj_dict(d)
        raise NotImplementedError(f'cannot parse CRS {crs!r} without pyproj')
    if isinstance(crs, dict):
        return _from_proj_dict(crs)
    # pyproj.CRS duck type
    if hasattr(crs, 'to_epsg'):
        epsg = crs.to_epsg()
        if epsg is not None:
            return parse_crs(int(epsg))
    if hasattr(crs, 'to_dict'):
        return _from_proj_dict(crs.to_dict())
    raise TypeError(f'cannot interpret {type(crs)} as a CRS')
