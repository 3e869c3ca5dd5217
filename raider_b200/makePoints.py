"""Drop-in for the reference's Cython module ``RAiDER.makePoints`` (tools/bindings/utils/makePoints.pyx:15-148).

``ray[..., c, k] = Rays_SP[..., c] + (k * stepSize) * Rays_SLV[..., c]`` with the reference's ``Npts`` rule,
generated on the device; output layout ``(..., 3, Npts)`` and float64 like the Cython original.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


def _make(max_len, Rays_SP, Rays_SLV, stepSize, ndim):
    for name, a in (('Rays_SP', Rays_SP), ('Rays_SLV', Rays_SLV)):
        if not isinstance(a, np.ndarray):
            raise TypeError(f"Argument '{name}' has incorrect type (expected numpy.ndarray, got {type(a).__name__})")
        if a.dtype != np.float64:
            raise ValueError(f"Buffer dtype mismatch, expected 'double' but got '{a.dtype}'")
        if a.ndim != ndim:
            raise ValueError(f'Buffer has wrong number of dimensions (expected {ndim}, got {a.ndim})')
    if Rays_SP.shape != Rays_SLV.shape or Rays_SP.shape[-1] != 3:
        raise ValueError('Rays_SP and Rays_SLV must have the same (..., 3) shape')
    lib = _lib.load()
    if float(stepSize) == 0.0:
        raise ZeroDivisionError('float modulo')
    npts = C.c_int64(0)
    check(lib.rdr_make_points_count(float(max_len), float(stepSize), C.byref(npts)))
    sp = np.ascontiguousarray(Rays_SP)
    slv = np.ascontiguousarray(Rays_SLV)
    n_rays = sp.size // 3
    out = np.empty(sp.shape + (max(npts.value, 0),), dtype=np.float64)
    if out.size:
        check(lib.rdr_make_points(float(max_len), ptr(sp), ptr(slv), n_rays, float(stepSize), ptr(out), npts.value,
                                  _lib.default_device(), _lib.MEM_HOST))
    return out


def makePoints0D(max_len, Rays_SP, Rays_SLV, stepSize):
    """makePoints.pyx:15-41: (3,) -> (3, Npts)."""
    return _make(max_len, Rays_SP, Rays_SLV, stepSize, 1)


def makePoints1D(max_len, Rays_SP, Rays_SLV, stepSize):
    """makePoints.pyx:45-74: (Nx, 3) -> (Nx, 3, Npts)."""
    return _make(max_len, Rays_SP, Rays_SLV, stepSize, 2)


def makePoints2D(max_len, Rays_SP, Rays_SLV, stepSize):
    """makePoints.pyx:79-110: (Nx, Ny, 3) -> (Nx, Ny, 3, Npts)."""
    return _make(max_len, Rays_SP, Rays_SLV, stepSize, 3)


def makePoints3D(max_len, Rays_SP, Rays_SLV, stepSize):
    """makePoints.pyx:115-148: (Nx, Ny, Nz, 3) -> (Nx, Ny, Nz, 3, Npts)."""
    return _make(max_len, Rays_SP, Rays_SLV, stepSize, 4)
