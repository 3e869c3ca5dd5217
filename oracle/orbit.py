"""CPU restatement of the orbit-based look vectors of the ray-tracing path -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference call sites (paths relative to /root/reference):

* ``Raytracing.getLookVectors``   tools/RAiDER/losreader.py:219-255 -- per pixel: ``isce3.geometry.geo2rdr(llh, ellipsoid,
  orbit, LUT2d() /* zero Doppler */, 0.06, look_dir, threshold=1e-7, maxiter=30, delta_range=10)``, then
  ``orbit.interpolate(aztime)`` and ``los = (sat_xyz - target_xyz) / slant_range``; any failure -> NaN vector.
* ``get_orbit``                   tools/RAiDER/losreader.py:736-769 -- state vectors sorted by time, duplicates dropped.

The arithmetic lives in a third-party dependency that is absent from /root/reference and not installable offline:
**isce3 (``isce3>=0.15.0``, environment.yml:25)**.  What is restated here is its published algorithm:

* ``isce3::core::Orbit::interpolate`` with the default ``OrbitInterpMethod::Hermite``
  (cxx/isce3/core/detail/InterpolateOrbit.icc): the 4-point osculating (Hermite) polynomial on uniformly spaced state
  vectors -- the ``orbitHermite`` routine inherited from ROI_PAC / ISCE2 -- through the state vectors ``i-2 .. i+1`` where
  ``i`` is the first vector with ``time[i] >= t`` (clamped to the ends);
* ``isce3::geometry::geo2rdr`` (cxx/isce3/geometry/geometry.cpp, Newton form): start at the orbit's mid time; each
  iteration interpolates the orbit, ``dr = target - sat``, stops when the slant range changed by less than ``threshold``,
  else ``aztime -= f / f'`` with ``f = dr . v - fdop * |dr|``, ``f' = -v . v + (fdop / |dr| + dfdop/dr) (dr . v)``
  (``fdop = 0`` for the zero-Doppler LUT the reference passes); not converged after ``maxiter`` -> failure.

PIN STATUS.  isce3 itself is not available offline, so there is no bit-for-bit pin of these two routines.  They are pinned
END TO END by the reference's own golden: ``test/test_slant.py:99`` (2.97711681 m: ``Raytracing`` on the Sentinel-1 precise-orbit
file ``test/orbit_files/S1B_OPER_AUX_POEORB_...EOF`` through the reference's ERA-5 cube) is reproduced to 5e-8 m -- inside the 7
decimals the reference asserts -- by the reference's own Python with THIS module standing in for isce3
(tests/test_oracle_vs_reference_py.py::test_reference_goldens_of_test_slant_are_reproduced, and from the committed fixture in
tests/test_oracle_pins.py).  A slant delay moves by about 2.5 m per radian of look direction at that geometry, so the golden
bounds the look-vector error of this restatement at ~2e-8 rad (centimetres of sensor position).  Below that, parity with isce3's
own rounding is unpinned.  Closed forms pin the pieces as well (tests/test_oracle_pins.py): a circular orbit (the construction of
test/fake_raytracing:73-117) where zero-Doppler time and slant range are known analytically, Hermite interpolation reproducing
its nodes and a degree-7 polynomial orbit exactly, and the zero-Doppler property ``(sat - target) . v_sat = 0`` on the reference's
Sentinel-1 state vectors (test/orbit_files/S1_sv_file.txt = test/test_losreader.py:20-92).
"""
from __future__ import annotations

import numpy as np

from . import geodesy


class Orbit:
    """Uniformly sampled state vectors (isce3.core.Orbit requires a Linspace of times): t [s], pos (n,3), vel (n,3)."""

    def __init__(self, t, pos, vel) -> None:
        t = np.asarray(t, dtype=np.float64)
        order = np.argsort(t, kind='stable')
        t, pos, vel = t[order], np.asarray(pos, dtype=np.float64)[order], np.asarray(vel, dtype=np.float64)[order]
        keep = np.concatenate([[True], np.diff(t) != 0])  # losreader.py:756-764: unique times
        self.t, self.pos, self.vel = t[keep], pos[keep], vel[keep]
        if self.t.size < 4:
            raise ValueError('at least 4 state vectors are required for Hermite interpolation')
        dt = np.diff(self.t)
        if not np.allclose(dt, dt[0], rtol=0, atol=1e-6 * abs(dt[0])):
            raise ValueError('state vectors must be uniformly spaced in time')
        self.spacing = (self.t[-1] - self.t[0]) / (self.t.size - 1)

    @property
    def mid_time(self) -> float:
        return float(self.t[0] + 0.5 * (self.t[-1] - self.t[0]))

    def interpolate(self, time: float):
        """(position, velocity) at ``time``; NaNs outside [t[0], t[-1]] (OrbitInterpBorderMode::FillNaN)."""
        if not (self.t[0] <= time <= self.t[-1]):
            return np.full(3, np.nan), np.full(3, np.nan)
        idx = int(np.argmax(self.t >= time)) - 2
        idx = min(max(idx, 0), self.t.size - 4)
        return orbit_hermite(self.pos[idx:idx + 4], self.vel[idx:idx + 4], self.t[idx:idx + 4], time)


def orbit_hermite(x, v, t, time):
    """ROI_PAC / ISCE ``orbitHermite`` on 4 state vectors: position and velocity of the degree-7 osculating polynomial."""
    f0, f1, h, hdot, g0, g1 = (np.zeros(4) for _ in range(6))
    for i in range(4):
        f1[i] = time - t[i]
        s = 0.0
        for j in range(4):
            if j != i:
                s += 1.0 / (t[i] - t[j])
        f0[i] = 1.0 - 2.0 * (time - t[i]) * s
    for i in range(4):
        product = 1.0
        for k in range(4):
            if k != i:
                product *= (time - t[k]) / (t[i] - t[k])
        h[i] = product
        s = 0.0
        for j in range(4):
            product = 1.0
            for k in range(4):
                if k != i and k != j:
                    product *= (time - t[k]) / (t[i] - t[k])
            if j != i:
                s += 1.0 / (t[i] - t[j]) * product
        hdot[i] = s
    for i in range(4):
        g1[i] = h[i] + 2.0 * (time - t[i]) * hdot[i]
        s = 0.0
        for j in range(4):
            if i != j:
                s += 1.0 / (t[i] - t[j])
        g0[i] = 2.0 * (f0[i] * hdot[i] - h[i] * s)
    pos = np.zeros(3)
    vel = np.zeros(3)
    for i in range(4):
        pos += (x[i] * f0[i] + v[i] * f1[i]) * h[i] * h[i]
        vel += (x[i] * g0[i] + v[i] * g1[i]) * h[i]
    return pos, vel


def geo2rdr(target_xyz, orbit: Orbit, threshold: float = 1.0e-7, maxiter: int = 30):
    """Zero-Doppler (aztime, slant_range) of an ECEF target; raises RuntimeError when Newton does not converge."""
    aztime = orbit.mid_time
    slant_old = 0.0
    for _ in range(maxiter):
        pos, vel = orbit.interpolate(aztime)
        dr = np.asarray(target_xyz, dtype=np.float64) - pos
        slant = float(np.sqrt(dr @ dr))
        if abs(slant - slant_old) < threshold:
            return aztime, slant
        slant_old = slant
        fn = float(dr @ vel)          # fdop == 0
        fnprime = -float(vel @ vel)
        aztime -= fn / fnprime
    raise RuntimeError('geo2rdr failed to converge')


def geo2rdr_llh(lon_rad: float, lat_rad: float, hgt: float, orbit: Orbit, threshold: float = 1.0e-7, maxiter: int = 30):
    """``isce3.geometry.geo2rdr(llh, ellipsoid, orbit, ...)`` as losreader.py:240-250 calls it: the target is given as
    (lon, lat) in RADIANS + height and converted to ECEF on the WGS84 ellipsoid (``Ellipsoid::lonLatToXyz``) before the solve."""
    xyz = np.array(geodesy.lla2ecef(np.rad2deg(lat_rad), np.rad2deg(lon_rad), hgt), dtype=np.float64)
    return geo2rdr(xyz, orbit, threshold=threshold, maxiter=maxiter)


class OrbitLOS:
    """LOS provider with the duck type of losreader.py:219-255: ECEF unit vectors ground -> sensor from an orbit."""

    def __init__(self, orbit: Orbit) -> None:
        self.orbit = orbit

    def getLookVectors(self, ht, llh, xyz, yy):
        yy = np.asarray(yy)
        los = np.full(yy.shape + (3,), np.nan)
        lon_r, lat_r = np.deg2rad(llh[0]), np.deg2rad(llh[1])    # losreader.py:226-228
        for ii in range(yy.shape[0]):
            for jj in range(yy.shape[1]):
                p = xyz[ii, jj, :]
                if np.isnan(p).any() or np.isnan(lon_r[ii, jj]) or np.isnan(lat_r[ii, jj]) or np.isnan(ht):
                    continue
                try:
                    aztime, slant = geo2rdr_llh(lon_r[ii, jj], lat_r[ii, jj], ht, self.orbit)   # the solve sees llh ...
                    sat, _ = self.orbit.interpolate(aztime)
                    los[ii, jj, :] = (sat - p) / slant                                         # ... the vector the caller's xyz
                except RuntimeError:
                    pass
        return los


def look_vectors_points(lat, lon, hgt, orbit: Orbit):
    """Look vectors for flat point lists (lat, lon in degrees)."""
    lat, lon, hgt = (np.atleast_1d(np.asarray(a, dtype=np.float64)) for a in (lat, lon, hgt))
    xyz = np.stack(geodesy.lla2ecef(lat, lon, hgt), axis=-1)
    los = np.full(lat.shape + (3,), np.nan)
    for i in range(lat.size):
        if np.isnan(xyz[i]).any():
            continue
        try:
            aztime, slant = geo2rdr(xyz[i], orbit)
            los[i] = (orbit.interpolate(aztime)[0] - xyz[i]) / slant
        except RuntimeError:
            pass
    return los
