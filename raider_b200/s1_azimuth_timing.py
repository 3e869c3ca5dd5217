"""Per-voxel temporal weights for the two / three weather-model epochs around an acquisition.

Reference: tools/RAiDER/s1_azimuth_timing.py -- ``get_azimuth_time_grid`` (:78-149: a Python triple loop over the voxels of the
cube calling isce3's ``geo2rdr``), ``get_inverse_weights_for_dates`` (:337-399) -- used by ``combine_weather_files`` with
``interp_method='azimuth_time_grid'`` (cli/raider.py:792-833): every voxel of the cube is weighted by the inverse of the time
between its own zero-Doppler (azimuth) time and the model epochs.

Here the azimuth-time grid comes from the device (K6, ``rdr_orbit_los``: the zero-Doppler Newton solve on Hermite-interpolated
state vectors, all voxels in one launch), the weights are the reference's arithmetic in NumPy, and the blend itself is
:meth:`raider_b200.engine.DeviceCube.from_epochs`.  Orbit *files* are the caller's business (the reference downloads them
through ASF / hyp3lib: out of scope); pass a :class:`raider_b200.losreader.Orbit`.
"""
from __future__ import annotations

import datetime as dt
from typing import Optional

import numpy as np

SPEED_OF_LIGHT = 299792458.0  # isce3.core.speed_of_light


def get_azimuth_time_grid(lon_mesh, lat_mesh, hgt_mesh, orb, device=None) -> np.ndarray:
    """Azimuth (zero-Doppler) time + range delay of every voxel, ``datetime64[ms]`` like the reference (:78-149); voxels whose
    solve does not converge / leave the orbit's span are ``NaT``."""
    from .losreader import orbit_look_vectors
    lon, lat, hgt = np.broadcast_arrays(np.asarray(lon_mesh, dtype=np.float64), np.asarray(lat_mesh, dtype=np.float64), np.asarray(hgt_mesh, dtype=np.float64))
    if orb.reference_epoch is None:
        raise ValueError('the orbit needs a reference epoch (build it from datetimes) to give absolute azimuth times')
    _, sr, az = orbit_look_vectors(orb, lat.ravel(), lon.ravel(), hgt.ravel(), threshold=1.0e-7, maxiter=100, device=device)
    t = az + sr / SPEED_OF_LIGHT          # seconds since the orbit's reference epoch (:141-143)
    out = np.full(t.shape, np.datetime64('NaT'), dtype='datetime64[ms]')
    ok = np.isfinite(t)
    epoch = np.datetime64(orb.reference_epoch, 'ms')
    out[ok] = epoch + (np.floor(t[ok] * 1e3)).astype('timedelta64[ms]')
    return out.reshape(lon.shape)


def get_s1_azimuth_time_grid(lon, lat, hgt, orb, device=None) -> np.ndarray:
    """``hgt x lat x lon`` azimuth-time cube from 1-D coordinate vectors or 3-D meshes (:152-214), for a given orbit."""
    lon, lat, hgt = np.asarray(lon), np.asarray(lat), np.asarray(hgt)
    dims = [c.ndim for c in (lon, lat, hgt)]
    if not all(d == dims[0] for d in dims):
        raise ValueError('All coordinates have same dimension (either 1 or 3 dimensional)')
    if dims[0] not in (1, 3):
        raise ValueError('Coordinates must be 1d or 3d coordinate arrays')
    if dims[0] == 1:
        hgt, lat, lon = np.meshgrid(hgt, lat, lon, indexing='ij')
    return get_azimuth_time_grid(lon, lat, hgt, orb, device=device)


def get_inverse_weights_for_dates(azimuth_time_array: np.ndarray, dates: list, inverse_regularizer: float = 1e-9,
                                  temporal_window_hours: Optional[float] = None) -> list:
    """Inverse-|time difference| weights per voxel and date, masked by the temporal window, normalised to sum 1 (:337-399)."""
    if len(set(dates)) != len(dates):
        raise ValueError('Dates provided must be unique')
    if not dates:
        raise ValueError('No dates provided')
    if not all(isinstance(d, dt.datetime) for d in dates):
        raise TypeError('dates must be all datetimes')
    if temporal_window_hours is None:
        window_s = min(abs((d - dates[0]).total_seconds()) for d in dates[1:])
    else:
        window_s = temporal_window_hours * 60 * 60
    diffs = [np.abs(azimuth_time_array - np.datetime64(d)) / np.timedelta64(1, 's') for d in dates]
    masked = [(1.0 / (diff + inverse_regularizer)) * (diff <= window_s).astype(int) for diff in diffs]
    if all((diff <= window_s).sum() == 0 for diff in diffs):
        raise ValueError('No dates provided are within temporal window')
    total = np.sum(np.stack(masked, axis=-1), axis=-1)
    return [m / total for m in masked]


def get_weights_time_interp(times: list, time: dt.datetime):
    """Scalar weights of the 'center_time' method (cli/raider.py:877-888)."""
    date1, date2 = times
    span = (date2 - date1).total_seconds()
    wgts = [1 - (time - date1).total_seconds() / span, 1 - (date2 - time).total_seconds() / span]
    return wgts if np.isclose(np.sum(wgts), 1) else None
