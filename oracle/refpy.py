"""Import the reference's OWN Python for the ray-trace path, unmodified -- TEST INFRASTRUCTURE (see oracle/__init__.py).

``RAiDER.delay`` / ``RAiDER.losreader`` / ``RAiDER.delayFcns`` / ``RAiDER.utilFcns`` under
``/root/reference/tools/RAiDER`` are pure NumPy + scipy on this path, but they ``import`` five packages that are
not installable offline: ``pyproj``, ``xarray``, ``rasterio``, ``isce3`` and ``shapely``.  This module installs
minimal stand-ins for those names in ``sys.modules`` and then imports the reference modules from where they lie
(nothing is copied).  What runs is the reference's own source for

* ``losreader.getTopOfAtmosphere`` (losreader.py:706-733), ``losreader.build_ray`` (:772-835),
  ``losreader.inc_hd_to_enu`` (:374-396), ``losreader.getZenithLookVecs`` (:302-316), the ``LOS`` classes (:32-299)
* ``delay._build_cube_ray`` (delay.py:219-326), ``delay._build_cube`` (:196-216), ``delay.transformPoints`` (:404-436)
* ``delayFcns.getInterpolators`` (delayFcns.py:23-58) on an in-memory ``xr.Dataset`` stand-in
* ``utilFcns.lla2ecef / ecef2lla / enu2ecef / ecef2enu`` (utilFcns.py:77-137)

What does NOT run is PROJ: the ``pyproj.Transformer`` stand-in performs the three conversions the path asks for
(4326<->4978 ``cart``, 4978->spherical ``lcc``, ``lcc``->4326) with :mod:`oracle.geodesy`, which restates PROJ's
published algorithms.  So this pins :mod:`oracle.raytrace` to the reference's *loop structure and NumPy arithmetic*
bit for bit; parity with PROJ's own rounding below ~1e-9 m stays unpinned (stated in DESIGN.md).

Only available where ``/root/reference`` exists (this container).  The GPU box has the golden vectors that
``tests/golden/make_golden.py`` produced with these functions.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

from . import geodesy

REFERENCE_ROOT = os.environ.get('RAIDER_REFERENCE_ROOT', '/root/reference')
_PKG_DIR = os.path.join(REFERENCE_ROOT, 'tools', 'RAiDER')


def available() -> bool:
    return os.path.isfile(os.path.join(_PKG_DIR, 'delay.py'))


# ------------------------------------------------------------------------------------------------------------------
# pyproj stand-in
# ------------------------------------------------------------------------------------------------------------------
class CRSError(Exception):
    pass


class CRS:
    """The slice of ``pyproj.CRS`` the path touches: construction from an EPSG code / dict / CRS, and ``==``."""

    def __init__(self, spec=None, **kw):
        if isinstance(spec, CRS):
            self.kind, self.params = spec.kind, dict(spec.params)
            return
        if spec is None and kw:
            spec = kw
        if isinstance(spec, str):
            s = spec.strip()
            if s.upper().startswith('EPSG:'):
                spec = int(s[5:])
            elif s.isdigit():
                spec = int(s)
            else:
                raise CRSError(f'unsupported CRS string for the stand-in: {spec!r}')
        if isinstance(spec, (int, np.integer)):
            code = int(spec)
            if code == 4326:
                self.kind, self.params = 'geographic', {}
            elif code == 4978:
                self.kind, self.params = 'geocentric', {}
            else:
                raise CRSError(f'EPSG:{code} is not known to the stand-in')
            return
        if isinstance(spec, dict):
            proj = spec.get('proj')
            if proj == 'lcc':
                a = spec.get('a', spec.get('R'))
                b = spec.get('b', a)
                if a is None or a != b:
                    raise CRSError('the stand-in implements the spherical lcc only')
                self.kind = 'lcc'
                self.params = dict(lat_1=float(spec['lat_1']), lat_2=float(spec['lat_2']), lat_0=float(spec['lat_0']),
                                   lon_0=float(spec['lon_0']), R=float(a),
                                   x_0=float(spec.get('x_0', 0.0)), y_0=float(spec.get('y_0', 0.0)))
                return
            if proj in ('longlat', 'latlong'):
                self.kind, self.params = 'geographic', {}
                return
        raise CRSError(f'unsupported CRS for the stand-in: {spec!r}')

    @classmethod
    def from_epsg(cls, code):
        try:
            return cls(int(code))
        except (TypeError, ValueError) as exc:
            raise CRSError(str(exc)) from exc

    @classmethod
    def from_user_input(cls, spec):
        return cls(spec)

    @classmethod
    def from_dict(cls, d):
        return cls(dict(d))

    @classmethod
    def from_wkt(cls, wkt):
        raise CRSError('WKT is not parsed by the stand-in')

    def to_epsg(self):
        return {'geographic': 4326, 'geocentric': 4978}.get(self.kind)

    def to_wkt(self):
        return f'STANDIN[{self.kind},{sorted(self.params.items())}]'

    @property
    def axis_info(self):   # utilFcns.transform_bbox (utilFcns.py:598) reads axis_info[0].unit_name
        return [types.SimpleNamespace(unit_name='degree' if self.kind == 'geographic' else 'metre')]

    def __eq__(self, other):
        if not isinstance(other, CRS):
            try:
                other = CRS(other)
            except CRSError:
                return False
        return self.kind == other.kind and self.params == other.params

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((self.kind, tuple(sorted(self.params.items()))))

    def __repr__(self):
        return f'<stand-in CRS {self.kind} {self.params}>'


class Transformer:
    """``pyproj.Transformer.from_crs(a, b, always_xy=True).transform(x, y, z)`` for the CRS pairs on the path."""

    def __init__(self, src: CRS, dst: CRS):
        self.src, self.dst = src, dst
        self._lcc_src = geodesy.LambertConformalSphere(**src.params) if src.kind == 'lcc' else None
        self._lcc_dst = geodesy.LambertConformalSphere(**dst.params) if dst.kind == 'lcc' else None

    @classmethod
    def from_crs(cls, crs_from, crs_to, always_xy=False, **kw):
        if not always_xy:
            raise CRSError('the stand-in implements always_xy=True only (every call site on the path uses it)')
        return cls(CRS(crs_from), CRS(crs_to))

    def transform(self, xx, yy, zz=None, **kw):
        s, d = self.src.kind, self.dst.kind
        xx = np.asarray(xx, dtype=np.float64)
        yy = np.asarray(yy, dtype=np.float64)
        if zz is not None:
            zz = np.asarray(zz, dtype=np.float64)
        if s == d and self.src.params == self.dst.params:
            return (xx, yy) if zz is None else (xx, yy, zz)
        # to geographic lon / lat / h first
        if s == 'geographic':
            lon, lat, h = xx, yy, zz
        elif s == 'geocentric':
            lon, lat, h = geodesy.ecef2lla(xx, yy, zz)
        elif s == 'lcc':
            lon, lat = self._lcc_src.inverse(xx, yy)
            h = zz
        else:
            raise CRSError(s)
        if d == 'geographic':
            out = (lon, lat, h)
        elif d == 'geocentric':
            out = geodesy.lla2ecef(lat, lon, h)
        elif d == 'lcc':
            X, Y = self._lcc_dst.forward(lon, lat)
            out = (X, Y, h)
        else:
            raise CRSError(d)
        return out[:2] if zz is None else out


class Proj:
    def __init__(self, *a, **kw):
        raise CRSError('pyproj.Proj is not on the delay path; the stand-in does not implement it')


# ------------------------------------------------------------------------------------------------------------------
# xarray stand-in: only what getInterpolators (delayFcns.py:31-41) reads
# ------------------------------------------------------------------------------------------------------------------
class Dataset:
    def __init__(self, variables: dict):
        self.variables = dict(variables)

    def __getitem__(self, k):
        return self.variables[k]


def _load_dataset(path, *a, **kw):
    raise ImportError('xarray is not installed: the stand-in Dataset is built in memory (oracle.refpy.dataset)')


def dataset(cube: dict) -> Dataset:
    """In-memory stand-in for the processed weather-model file: {x, y, z, wet, hydro[, wet_total, hydro_total]}."""
    return Dataset({k: np.asarray(v) for k, v in cube.items()})


# ------------------------------------------------------------------------------------------------------------------
# isce3 stand-in: the six names losreader.py uses (losreader.py:24,193-198,240-251,593-598,656-686,751-767), backed by
# the restatement of isce3's published algorithms in oracle/orbit.py.  With it the reference's own Raytracing /
# Conventional classes, get_orbit and state_to_los run unmodified; what that pins is the CALL STRUCTURE around isce3.
# isce3's arithmetic itself is pinned only through the reference's end-to-end golden (test/test_slant.py:99), which
# tests/test_oracle_vs_reference_py.py reproduces with this stand-in.
# ------------------------------------------------------------------------------------------------------------------
class _IsceDateTime:
    def __init__(self, t):
        self.t = t.t if isinstance(t, _IsceDateTime) else t

    def _key(self):
        return self.t

    def __eq__(self, other):
        return isinstance(other, _IsceDateTime) and self.t == other.t

    def __lt__(self, other):
        return self.t < other.t

    def __hash__(self):
        return hash(self.t)


class _IsceStateVector:
    def __init__(self, datetime, position, velocity):
        self.datetime = _IsceDateTime(datetime)
        self.position = np.asarray(position, dtype=np.float64)
        self.velocity = np.asarray(velocity, dtype=np.float64)


class _IsceOrbit:
    """isce3.core.Orbit: the reference epoch is the first state vector's time; times are seconds since it."""

    def __init__(self, svs):
        from . import orbit as _orbit
        self.reference_epoch = svs[0].datetime
        t = np.array([(sv.datetime.t - self.reference_epoch.t).total_seconds() for sv in svs], dtype=np.float64)
        if np.any(np.diff(t) <= 0):
            raise ValueError('isce3.core.Orbit needs uniformly spaced, increasing state vectors')
        self._o = _orbit.Orbit(t, np.stack([sv.position for sv in svs]), np.stack([sv.velocity for sv in svs]))

    @property
    def time(self):
        return self._o.t

    @property
    def position(self):
        return self._o.pos

    @property
    def velocity(self):
        return self._o.vel

    def interpolate(self, t):
        pos, vel = self._o.interpolate(float(t))
        if np.isnan(pos).any():
            raise ValueError('orbit interpolation outside the state vectors')  # OrbitInterpBorderMode::Error (default)
        return pos, vel


class _IsceEllipsoid:
    a, e2 = geodesy.WGS84_A, geodesy.WGS84_ES

    @staticmethod
    def n_vector(lon, lat):
        return np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])


def _isce_geo2rdr(llh, ellipsoid, orbit, doppler, wavelength, side, threshold=1.0e-8, maxiter=50, delta_range=10.0):
    from . import orbit as _orbit
    lon, lat, h = (float(v) for v in llh)
    return _orbit.geo2rdr_llh(lon, lat, h, orbit._o, threshold=threshold, maxiter=maxiter)


def _isce_modules():
    look = types.SimpleNamespace(Right='right', Left='left')
    core = _module('isce3.ext.isce3.core', Orbit=_IsceOrbit, StateVector=_IsceStateVector, DateTime=_IsceDateTime,
                   Ellipsoid=_IsceEllipsoid, LUT2d=type('LUT2d', (), {}), LookSide=look)
    geometry = _module('isce3.ext.isce3.geometry', geo2rdr=_isce_geo2rdr)
    inner = _module('isce3.ext.isce3', core=core, geometry=geometry)
    ext = _module('isce3.ext', isce3=inner)
    top = _module('isce3', ext=ext, core=core, geometry=geometry)
    for m in (top, ext, inner):
        m.__path__ = []
    return {'isce3': top, 'isce3.ext': ext, 'isce3.ext.isce3': inner, 'isce3.ext.isce3.core': core,
            'isce3.ext.isce3.geometry': geometry}


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__raider_b200_standin__ = True
    return m


_STANDINS = ('pyproj', 'pyproj.exceptions', 'xarray', 'rasterio', 'shapely', 'shapely.geometry')
_loaded = None


def load():
    """Import the reference modules; returns a namespace with ``delay``, ``losreader``, ``delayFcns``, ``utilFcns``."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError(f'{_PKG_DIR} not found: the reference tree exists only in the build container')

    exc_mod = _module('pyproj.exceptions', CRSError=CRSError)
    installs = {
        'pyproj': _module('pyproj', CRS=CRS, Transformer=Transformer, Proj=Proj, exceptions=exc_mod),
        'pyproj.exceptions': exc_mod,
        'xarray': _module('xarray', Dataset=Dataset, DataArray=type('DataArray', (), {}), load_dataset=_load_dataset,
                          open_dataset=_load_dataset),
        'rasterio': _module('rasterio', crs=_module('rasterio.crs', CRS=type('CRS', (), {})),
                            transform=_module('rasterio.transform', Affine=type('Affine', (), {}))),
        'shapely': _module('shapely'),
        'shapely.geometry': _module('shapely.geometry', Polygon=object, Point=object, box=None),
    }
    installs.update(_isce_modules())
    installs['rasterio.crs'] = installs['rasterio'].crs
    installs['rasterio.transform'] = installs['rasterio'].transform
    installs['rasterio'].__path__ = []
    for name, mod in installs.items():
        if name not in sys.modules:
            sys.modules[name] = mod

    if 'RAiDER' not in sys.modules:
        # RAiDER/__init__.py asks importlib.metadata for the installed version; the package is not installed, so the
        # package object is created here and its submodules are imported from the reference tree as they are.
        pkg = types.ModuleType('RAiDER')
        pkg.__path__ = [_PKG_DIR]
        pkg.__version__ = '0+reference'
        sys.modules['RAiDER'] = pkg

    # RAiDER.logger opens debug.log / error.log in conf.LOGGER_PATH (default: the cwd) at import: point it at a scratch dir
    import logging
    import tempfile
    from pathlib import Path
    conf = importlib.import_module('RAiDER.cli.conf')
    conf.setLoggerPath(Path(tempfile.mkdtemp(prefix='raider_ref_log_')))
    ref_logger = importlib.import_module('RAiDER.logger').logger
    for h in ref_logger.handlers:
        if isinstance(h, logging.StreamHandler) and not isinstance(h, logging.FileHandler):
            h.setLevel(logging.WARNING)

    ns = types.SimpleNamespace()
    for name in ('constants', 'utilFcns', 'losreader', 'delayFcns', 'delay'):
        setattr(ns, name, importlib.import_module(f'RAiDER.{name}'))
    # the inverse-time weights of the azimuth_time_grid interpolation (s1_azimuth_timing.py:337-399) are pure NumPy; the module's
    # other imports (ASF search, orbit download) get empty stand-ins
    for name, mod in {'asf_search': _module('asf_search'), 's1_orbits': _module('s1_orbits')}.items():
        sys.modules.setdefault(name, mod)
    shp = sys.modules['shapely.geometry']
    if getattr(shp, '__raider_b200_standin__', False) and getattr(shp, 'Point', object) is object:
        shp.Point = type('Point', (), {})
    try:
        ns.s1_azimuth_timing = importlib.import_module('RAiDER.s1_azimuth_timing')
    except Exception:   # optional: only its weights function is used, by one test
        ns.s1_azimuth_timing = None
    ns.CRS = CRS
    ns.Transformer = Transformer
    ns.dataset = dataset
    _loaded = ns
    return ns


# ------------------------------------------------------------------------------------------------------------------
# LOS providers for the reference's duck type (delay.py:270) where the reference class needs isce3 / files
# ------------------------------------------------------------------------------------------------------------------
def fixed_incidence_los(ref, incidence_deg: float, heading_deg: float):
    """Constant incidence/heading through the reference's inc_hd_to_enu (losreader.py:374-396) + enu2ecef
    (utilFcns.py:91-121) -- the reference has no ray-tracing class for this, SURVEY a12."""
    class _Fixed:
        def getLookVectors(self, ht, llh, xyz, yy):
            enu = ref.losreader.inc_hd_to_enu(np.float64(incidence_deg), np.float64(heading_deg))
            return ref.utilFcns.enu2ecef(enu[..., 0], enu[..., 1], enu[..., 2], llh[1], llh[0], llh[2])
    return _Fixed()


def zenith_los(ref):
    class _Zen:
        def getLookVectors(self, ht, llh, xyz, yy):
            return ref.losreader.getZenithLookVecs(llh[1], llh[0], llh[2])
    return _Zen()


def array_los(vecs):
    class _Arr:
        def getLookVectors(self, ht, llh, xyz, yy):
            return np.asarray(vecs, dtype=np.float64)
    return _Arr()
