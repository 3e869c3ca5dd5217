"""Drop-in for the reference's pybind11 module ``RAiDER.interpolate`` (tools/bindings/interpolate/src/module.cpp).

Same two functions, same argument names/defaults, same error conventions (``TypeError`` for shape/dimension/axis
problems, ``RuntimeError`` for axis 0 with several threads), evaluated by CUDA kernels through the C ABI.
``assume_sorted`` and ``max_threads`` are accepted for signature compatibility: the former only selects a search
strategy in the reference (identical results for inputs that honour the promise), the latter has no meaning on a GPU
except for the reference's axis-0 check, which is kept.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


def _c64(a):
    return np.asarray(a, dtype=np.float64, order='C')  # (ascontiguousarray would promote 0-d to 1-d)


def interpolate(points, values, interp_points, fill_value=None, assume_sorted=False, max_threads=8):
    """Linear interpolator in any dimension (module.cpp:26-294); ndim <= 8 on the device."""
    points = [_c64(p) for p in points]
    values = _c64(values)
    interp_points = _c64(interp_points)
    num_dims = len(points)
    if values.ndim == 0 or interp_points.ndim == 0:
        raise TypeError('Only arrays are supported, not scalar values!')
    for arr in points:
        if arr.ndim != 1:
            raise TypeError("'points' must be a list of 1D arrays!")
    if num_dims != values.ndim:
        raise TypeError(f'Dimension mismatch! Grid is {num_dims}D but values are {values.ndim}D!')
    if interp_points.ndim != 2:
        raise TypeError("'interp_points' should have shape (N, ndim).")
    if interp_points.shape[1] != num_dims:
        raise TypeError(f'Dimension mismatch! Grid is {num_dims}D but interpolation points are {interp_points.shape[1]}D!')
    for d, p in enumerate(points):
        if p.size != values.shape[d]:
            raise TypeError(f'Dimension mismatch! Grid axis {d} has {p.size} nodes but values has {values.shape[d]}!')
    n = interp_points.shape[0]
    out = np.empty(n, dtype=np.float64)
    grids = (C.c_void_p * num_dims)(*[p.ctypes.data for p in points])
    sizes = (C.c_int64 * num_dims)(*[p.size for p in points])
    has_fill = fill_value is not None
    check(_lib.load().rdr_interpolate(num_dims, grids, sizes, ptr(values), ptr(interp_points), n, int(has_fill),
                                      float(fill_value) if has_fill else 0.0, ptr(out), _lib.default_device(), _lib.MEM_HOST))
    return out


def interpolate_along_axis(points, values, interp_points, axis=-1, fill_value=None, assume_sorted=False, max_threads=8):
    """1D linear interpolator along a specific axis (module.cpp:296-493)."""
    points = _c64(points)
    values = _c64(values)
    interp_points = _c64(interp_points)
    if values.ndim == 0 or interp_points.ndim == 0:
        raise TypeError('Only arrays are supported, not scalar values!')
    if points.ndim != values.ndim or points.ndim != interp_points.ndim:
        raise TypeError("'points', 'values' and 'interp_points' must all have the same number of dimensions!")
    dimensions = points.ndim
    for i in range(dimensions):
        if points.shape[i] != values.shape[i]:
            raise TypeError("'points' and 'values' must have the same shape!")
    if axis < 0:
        axis += dimensions
    if axis >= dimensions or axis < 0:
        raise TypeError("'axis' out of range!")
    elif axis == 0 and max_threads > 1:
        raise RuntimeError('Cannot interpolate along axis 0 with multiple threads!')
    for i in range(dimensions):
        if i != axis and interp_points.shape[i] != points.shape[i]:
            raise TypeError(
                f"Dimension mismatch at axis {i}! 'points' is {points.shape[i]} but interp_points is {interp_points.shape[i]}!"
            )
    nin, nout = points.shape[axis], interp_points.shape[axis]
    x = np.ascontiguousarray(np.moveaxis(points, axis, -1)).reshape(-1, nin)
    y = np.ascontiguousarray(np.moveaxis(values, axis, -1)).reshape(-1, nin)
    qm = np.moveaxis(interp_points, axis, -1)
    q = np.ascontiguousarray(qm).reshape(-1, nout)
    out = np.empty_like(q)
    has_fill = fill_value is not None
    check(_lib.load().rdr_interp_along_axis(ptr(x), ptr(y), ptr(q), x.shape[0], nin, nout, int(has_fill),
                                            float(fill_value) if has_fill else 0.0, ptr(out), _lib.default_device(), _lib.MEM_HOST))
    return np.ascontiguousarray(np.moveaxis(out.reshape(qm.shape), -1, axis))
