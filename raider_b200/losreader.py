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
import datetime as dt
import os
import xml.etree.ElementTree as ET
from pathlib import PosixPath
from typing import Union

import numpy as np

from . import _lib
from ._lib import check, f64, ptr
from .constants import _ZREF
from .utilFcns import cosd, enu2ecef, sind


def _reference_los_base():
    """RAiDER's own ``LOS`` base when a real RAiDER install is importable: the LOS classes *remain* RAiDER's
    (BASELINE.json north_star), so objects built here pass ``isinstance(los, RAiDER.losreader.LOS)`` there."""
    try:
        import importlib.util
        if importlib.util.find_spec('RAiDER') is None:
            return None
        from RAiDER.losreader import LOS as reference_los
        return reference_los
    except Exception:
        return None


class _LOSContract:
    """The duck type ``RAiDER.delay`` relies on (contract of losreader.py:32-72), stated minimally: three mode
    predicates, ``setTime``, and ``setPoints`` accepting (lats, lons, heights), (lats, lons) or one (..., 3) array."""

    _ray_trace = _is_zenith = _is_projected = False
    _lats = _lons = _heights = _look_vecs = _time = None

    def __init__(self) -> None:
        pass

    def setPoints(self, lats, lons=None, heights=None) -> None:
        if lats is None and self._lats is None:
            raise RuntimeError("You haven't given any point locations yet")
        if lons is None:  # one array of [lat, lon, height] triples
            lats, lons, heights = (lats[..., c] for c in range(3))
        elif heights is None:
            heights = np.zeros((len(lats), 1))
        self._lats, self._lons, self._heights = lats, lons, heights

    def setTime(self, datetime) -> None:
        self._time = datetime

    def is_Zenith(self):
        return self._is_zenith

    def is_Projected(self):
        return self._is_projected

    def ray_trace(self):
        return self._ray_trace


LOS = _reference_los_base() or _LOSContract


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
            # orbit file: cos(look angle) per point from the zero-Doppler geometry (losreader.py:124-133 -> state_to_los)
            svs = np.stack(get_sv(self._file, self._time, self._pad), axis=-1)
            los_factor = state_to_los(svs, [np.asarray(self._lats), np.asarray(self._lons), np.asarray(self._heights)])
            return np.asarray(delays) / los_factor
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
    * ``Raytracing(filename=orbit_file, time=t)``: orbit state vectors (ESA .EOF, 7-column text; :478-518, :736-769);
      the zero-Doppler look vector of every pixel is solved on the device (K6) -- the reference loops over pixels in
      Python calling isce3's ``geo2rdr`` (:230-254).
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
        self._look_dir = look_dir.lower()
        self._incidence, self._heading, self._vecs = incidence, heading, look_vecs
        self._orbit = None
        if incidence is None and look_vecs is None and filename is None:
            raise ValueError('Raytracing needs an orbit file (filename=), incidence=/heading=, or look_vecs=')
        if incidence is not None:
            self._enu = inc_hd_to_enu(np.float64(incidence), np.float64(0.0 if heading is None else heading))
        elif look_vecs is None and self._time is not None:
            self._orbit = filename if isinstance(filename, Orbit) else get_orbit(self._file, self._time, pad=pad)
        elif look_vecs is None and isinstance(filename, Orbit):
            self._orbit = filename

    def getSensorDirection(self):
        """'desc' or 'asc' from the z component of the first / last state vector (losreader.py:198-205)."""
        if self._orbit is None:
            raise ValueError('The orbit has not been set')
        z, t = self._orbit.position[:, 2], self._orbit.time
        return 'desc' if z[np.argmin(t)] > z[np.argmax(t)] else 'asc'

    def getLookDirection(self):
        return self._look_dir

    def setTime(self, time, pad=600) -> None:
        """Called in checkArgs (losreader.py:210-212)."""
        self._time = time
        if self._incidence is None and self._vecs is None and not isinstance(self._file, Orbit):
            self._orbit = get_orbit(self._file, self._time, pad=pad)

    def device_spec(self):
        if self._incidence is not None and np.ndim(self._incidence) == 0:
            return _lib.LOS_ENU_CONST, f64(self._enu)
        if self._incidence is None and self._vecs is None:
            if self._orbit is None:
                raise ValueError('The orbit has not been set (pass time= or call setTime)')
            return _lib.LOS_ORBIT, self._orbit.packed()
        return None

    def getLookVectors(self, ht, llh, xyz, yy):
        """(ny, nx, 3) ECEF unit vectors ground -> sensor (contract of losreader.py:219-255)."""
        if self._vecs is not None:
            v = self._vecs(ht, llh, xyz, yy) if callable(self._vecs) else self._vecs
            return np.asarray(v, dtype=np.float64)
        if self._incidence is not None:
            e, n, u = self._enu
            return enu2ecef(e, n, u, llh[1], llh[0], llh[2])
        if self._orbit is None:
            raise ValueError('The orbit has not been set (pass time= or call setTime)')
        shape = np.shape(yy)
        lon = f64(np.broadcast_to(llh[0], shape)).ravel()
        lat = f64(np.broadcast_to(llh[1], shape)).ravel()
        hgt = f64(np.broadcast_to(np.asarray(llh[2], dtype=np.float64), shape)).ravel()
        los, _, _ = orbit_look_vectors(self._orbit, lat, lon, hgt)
        return los.reshape(shape + (3,))


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


class Orbit:
    """Time-ordered, unique, uniformly spaced state vectors -- the role of ``isce3.core.Orbit`` in losreader.py:736-769.

    ``time`` is seconds since ``reference_epoch`` (the first state vector), ``position`` / ``velocity`` are (n, 3) ECEF.
    """

    def __init__(self, times, position, velocity, reference_epoch=None) -> None:
        times = np.asarray(times)
        if times.dtype == object or np.issubdtype(times.dtype, np.datetime64):
            tlist = [t if isinstance(t, dt.datetime) else t.astype('datetime64[us]').astype(dt.datetime) for t in times]
            reference_epoch = reference_epoch or min(tlist)
            times = np.array([(t - reference_epoch).total_seconds() for t in tlist])
        times = np.asarray(times, dtype=np.float64)
        position, velocity = np.asarray(position, dtype=np.float64).reshape(-1, 3), np.asarray(velocity, dtype=np.float64).reshape(-1, 3)
        order = np.argsort(times, kind='stable')
        times, position, velocity = times[order], position[order], velocity[order]
        keep = np.concatenate([[True], np.diff(times) != 0])  # only unique state vectors (losreader.py:756-764)
        self.time, self.position, self.velocity = times[keep], position[keep], velocity[keep]
        self.reference_epoch = reference_epoch
        if self.time.size < 4:
            raise ValueError('Orbit: at least 4 state vectors are required for orbit interpolation')
        d = np.diff(self.time)
        if not np.allclose(d, d[0], rtol=0.0, atol=1e-6 * abs(d[0])):
            raise ValueError('Orbit: state vectors must be uniformly spaced in time')

    @property
    def size(self) -> int:
        return int(self.time.size)

    def packed(self) -> np.ndarray:
        """{n_sv, rows of (t, x, y, z, vx, vy, vz)}: the RDR_LOS_ORBIT payload of include/raider_b200.h."""
        rows = np.concatenate([self.time[:, None], self.position, self.velocity], axis=1)
        return np.ascontiguousarray(np.concatenate([[float(self.size)], rows.ravel()]))


def orbit_look_vectors(orbit: 'Orbit', lats, lons, heights, threshold=1.0e-7, maxiter=30, device=None):
    """Zero-Doppler look vectors, slant ranges and azimuth times of points (deg, deg, m) on the device (K6)."""
    lat, lon, hgt = f64(lats).ravel(), f64(lons).ravel(), f64(heights).ravel()
    n = lat.size
    los, sr, az = np.empty((n, 3)), np.empty(n), np.empty(n)
    rows = orbit.packed()
    if n:
        check(_lib.load().rdr_orbit_los(ptr(rows[1:]), orbit.size, _lib.GEOM_POINTS, ptr(lon), ptr(lat), ptr(hgt), 0.0, 1, n, float(threshold),
                                        int(maxiter), ptr(los), ptr(sr), ptr(az), _lib.default_device() if device is None else device))
    return los, sr, az


def read_txt_file(filename):
    """7-column text file of orbit state vectors: ISO time, x y z, vx vy vz (losreader.py:429-475)."""
    t, cols = [], [[] for _ in range(6)]
    with open(filename) as f:
        for line in f:
            try:
                parts = line.strip().split()
                t_ = dt.datetime.fromisoformat(parts[0])
                vals = [float(v) for v in parts[1:]]
                if len(vals) != 6:
                    raise ValueError
            except (ValueError, IndexError):
                raise ValueError(f'I need {filename} to be a 7 column text file, with columns t, x, y, z, vx, vy, vz '
                                 f"(Couldn't parse line {repr(line)})")
            t.append(t_)
            for c, v in zip(cols, vals):
                c.append(v)
    if len(t) < 4:
        raise ValueError(f'read_txt_file: File {filename} does not have enough statevectors')
    return [np.array(a) for a in [t] + cols]


def read_ESA_Orbit_file(filename):
    """Orbit state vectors of an ESA .EOF file: [t (datetimes), x, y, z, vx, vy, vz] (losreader.py:478-518)."""
    root = ET.parse(filename).getroot()
    osvs = root[1][0]
    n = len(osvs)
    t, arr = [], np.ones((6, n))
    for i, st in enumerate(osvs):
        t.append(dt.datetime.strptime(st[1].text, 'UTC=%Y-%m-%dT%H:%M:%S.%f'))
        for c in range(6):
            arr[c, i] = float(st[4 + c].text)
    return [np.array(t)] + [arr[c] for c in range(6)]


def filter_ESA_orbit_file(orbit_xml: str, ref_time: dt.datetime) -> bool:
    """True when the validity window in the .EOF file name contains ``ref_time`` (losreader.py:537-555)."""
    f = os.path.basename(orbit_xml)
    t0 = dt.datetime.strptime(f.split('_')[6].lstrip('V'), '%Y%m%dT%H%M%S')
    t1 = dt.datetime.strptime(f.split('_')[7].rstrip('.EOF'), '%Y%m%dT%H%M%S')
    return t0 < ref_time < t1


def pick_ESA_orbit_file(list_files: list, ref_time: dt.datetime):
    """From a list of .EOF orbit files, pick the one that contains ``ref_time`` (losreader.py:520-534)."""
    for path in list_files:
        if filter_ESA_orbit_file(path, ref_time):
            return path
    raise AssertionError('Given orbit files did not match given date/time')


def cut_times(times, ref_time, pad):
    """Mask of the orbit times within ``pad`` seconds of ``ref_time`` (losreader.py:609-628)."""
    diff = np.array([(x - ref_time).total_seconds() for x in times])
    return np.abs(diff) < pad


def get_sv(los_file: Union[str, list, PosixPath], ref_time: dt.datetime, pad: int):
    """State vectors of a text / ESA orbit file (or list of ESA files) around ``ref_time`` (losreader.py:319-371)."""
    try:
        svs = read_txt_file(los_file)
    except (ValueError, TypeError, UnicodeDecodeError, OSError):
        try:
            los_files = [los_file] if isinstance(los_file, (str, PosixPath)) else los_file
            los_files = sorted(list(set(los_files)))
            los_files = [p for p in los_files if filter_ESA_orbit_file(str(p), ref_time)]
            if not los_files:
                raise ValueError('There are no valid orbit files provided')
            svs = []
            for orb_path in los_files:
                svs.extend(read_ESA_Orbit_file(orb_path))
            if len(los_files) > 1:  # concatenate the per-file lists column by column
                svs = [np.concatenate(svs[c::7]) for c in range(7)]
        except Exception:
            raise ValueError(f'get_sv: I cannot parse the statevector file {los_file}')
    if ref_time:
        idx = cut_times(svs[0], ref_time, pad=pad)
        svs = [d[idx] for d in svs]
    return svs


def get_orbit(orbit_file: Union[list, str], ref_time: dt.datetime, pad: int) -> Orbit:
    """State vectors around ``ref_time``, unique and ordered in time (losreader.py:736-769)."""
    svs = get_sv(orbit_file, ref_time, pad)
    return Orbit(svs[0], np.stack(svs[1:4], axis=-1), np.stack(svs[4:7], axis=-1))


def get_radar_pos(llh, orb: Orbit, device=None):
    """Look angle (deg, between the target->sensor vector and the ellipsoid normal) and slant range (losreader.py:631-703)."""
    llh = np.asarray(llh, dtype=np.float64).reshape(-1, 3)
    lat, lon, hgt = llh[:, 0], llh[:, 1], llh[:, 2]
    los, sr, _ = orbit_look_vectors(orb, lat, lon, hgt, device=device)
    nv = getZenithLookVecs(lat, lon, hgt)  # isce3 Ellipsoid.n_vector: the ellipsoid normal
    cosang = np.clip(np.einsum('ij,ij->i', los, nv), -1.0, 1.0)
    return np.rad2deg(np.arccos(cosang)), sr


def state_to_los(svs, llh_targets):
    """cos(look angle) at every target from orbit state vectors, rows (t, x, y, z, vx, vy, vz) (losreader.py:558-606)."""
    svs = np.asarray(svs)
    if np.min(svs.shape) < 4:
        raise RuntimeError('state_to_los: At least 4 state vectors are required for orbit interpolation')
    orb = Orbit(svs[:, 0], svs[:, 1:4].astype(np.float64), svs[:, 4:7].astype(np.float64))
    in_shape = np.shape(llh_targets[0])
    target_llh = np.stack([np.asarray(x, dtype=np.float64).flatten() for x in llh_targets], axis=-1)
    los_ang, _ = get_radar_pos(target_llh, orb)
    return np.cos(np.deg2rad(los_ang)).reshape(in_shape)


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
