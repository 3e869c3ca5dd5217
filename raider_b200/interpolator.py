"""Python interpolation wrappers (reference: tools/RAiDER/interpolator.py:19-107)."""
from __future__ import annotations

from typing import Tuple, Union

import numpy as np

from .interpolate import interpolate, interpolate_along_axis


class RegularGridInterpolator:
    """Wrapper around ``interpolate`` with a scipy-like interface (interpolator.py:19-69)."""

    def __init__(self, grid, values, fill_value=None, assume_sorted: bool = False, max_threads: int = 8) -> None:
        self.grid = grid
        self.values = values
        self.fill_value = fill_value
        self.assume_sorted = assume_sorted
        self.max_threads = max_threads

    def __call__(self, points: Union[Tuple, np.ndarray]) -> np.ndarray:
        if isinstance(points, tuple):
            shape = points[0].shape
            for arr in points:
                assert arr.shape == shape, 'All dimensions must contain the same number of points!'
            interp_points = np.stack(points, axis=-1)
            in_shape = interp_points.shape
        elif points.ndim > 2:
            in_shape = points.shape
            interp_points = points.reshape((np.prod(points.shape[:-1]),) + (points.shape[-1],))
        else:
            interp_points = points
            in_shape = interp_points.shape

        out = interpolate(
            self.grid,
            self.values,
            interp_points,
            fill_value=self.fill_value,
            assume_sorted=self.assume_sorted,
            max_threads=self.max_threads,
        )
        return out.reshape(in_shape[:-1])


def interp_along_axis(oldCoord, newCoord, data, axis=2, pad=False):
    """DEPRECATED in the reference (interpolator.py:72-89) in favour of ``interpolate_along_axis``; same results.

    1-D ``oldCoord``/``newCoord`` are broadcast along the other axes; out-of-range points are NaN.
    """
    oldCoord = np.asarray(oldCoord, dtype=np.float64)
    newCoord = np.asarray(newCoord, dtype=np.float64)
    data = np.asarray(data, dtype=np.float64)
    if oldCoord.ndim == 1 and data.ndim > 1:
        shape = [1] * data.ndim
        shape[axis] = oldCoord.size
        oldCoord = np.broadcast_to(oldCoord.reshape(shape), data.shape)
        nshape = list(data.shape)
        nshape[axis] = newCoord.size
        shape[axis] = newCoord.size
        newCoord = np.broadcast_to(newCoord.reshape(shape), nshape)
    out = interpolate_along_axis(oldCoord, data, newCoord, axis=axis, fill_value=np.nan, max_threads=1)
    # np.interp / interp1d (what the reference version calls) treat the last node as in-bounds; bisect_left does not
    last_x = np.take(oldCoord, [-1], axis=axis)
    last_y = np.take(data, [-1], axis=axis)
    return np.where(newCoord == last_x, last_y, out)


def interpV(y, old_x, new_x, left=None, right=None, period=None):
    """Rearrange np.interp's arguments (interpolator.py:92-94)."""
    return np.interp(new_x, old_x, y, left=left, right=right, period=period)


def interpVector(vec, Nx):
    """interpolator.py:97-107: a single vector holding x, y and the new x, in that order."""
    vec = np.asarray(vec, dtype=np.float64)
    x = vec[:Nx]
    y = vec[Nx:2 * Nx]
    xnew = vec[2 * Nx:]
    return interpolate_along_axis(x, y, xnew, axis=0, fill_value=np.nan, max_threads=1)
