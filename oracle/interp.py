"""NumPy restatement of RAiDER's native numerics -- TEST INFRASTRUCTURE (see oracle/__init__.py).

* :func:`bisect_left`, :func:`find_left`     tools/bindings/interpolate/src/interpolate.h:23-56
* :func:`interpolate`                        interpolate.h:78-118 (1-D), interpolate.cpp:18-83 (2-D),
                                             :85-176 (3-D), :178-258 (N-D); binding module.cpp:26-294
* :func:`interpolate_along_axis`             interpolate.cpp:260-332; binding module.cpp:296-493
* :func:`makePoints`                         tools/bindings/utils/makePoints.pyx:15-148
* :func:`scipy_find_interval`                the interval rule of scipy's RegularGridInterpolator
                                             (the interpolator the delay path really uses, delayFcns.py:55-56)

The floating-point expressions keep the reference's operation order (the reference is built
without -march/-mfma, setup.py:31-37, i.e. no FMA contraction), so results are bit-identical to
the compiled reference in ``oracle/_ref`` -- tests/test_oracle_pins.py checks exactly that.
Error conventions follow module.cpp (TypeError / RuntimeError).
"""
from __future__ import annotations

import numpy as np


def bisect_left(grid, x):
    """interpolate.h:23-38: first index i with x < grid[i] (N if none; NaN compares false -> N)."""
    grid = np.asarray(grid, dtype=np.float64)
    left, right = 0, grid.shape[0]
    while right != left:
        mid = (left + right) // 2
        if x < grid[mid]:
            right = mid
        else:
            left = mid + 1
    return right


def find_left(grid, x):
    """interpolate.h:44-56: 5-step linear scan then bisection."""
    grid = np.asarray(grid, dtype=np.float64)
    n = grid.shape[0]
    left = 0
    while left < 5:
        if left == n or x < grid[left]:
            return left
        left += 1
    return bisect_left(grid[left:], x) + left


def _hi_index(grid, x, fill):
    """Vectorised bisect_left + fill/clamp decision (interpolate.cpp:106-131). Returns (hi, filled_mask)."""
    n = grid.shape[0]
    hi = np.searchsorted(grid, x, side='right')  # == bisect_left for every x incl. NaN (-> n)
    if fill:
        bad = (hi < 1) | (hi > n - 1)
        hi = np.clip(hi, 1, n - 1)
        return hi, bad
    return np.clip(hi, 1, n - 1), np.zeros(x.shape, dtype=bool)


def interpolate(points, values, interp_points, fill_value=None, assume_sorted=False, max_threads=8):
    """RAiDER.interpolate.interpolate (module.cpp:26-294).

    ``assume_sorted`` only changes *how* the interval is searched (find_left from the previous
    interval); for inputs that honour the promise the result is identical, so it is ignored here.
    """
    points = [np.asarray(p, dtype=np.float64, order='C') for p in points]
    values = np.asarray(values, dtype=np.float64, order='C')
    interp_points = np.asarray(interp_points, dtype=np.float64, order='C')
    ndim = len(points)
    if values.ndim == 0 or interp_points.ndim == 0:
        raise TypeError('Only arrays are supported, not scalar values!')
    for p in points:
        if p.ndim != 1:
            raise TypeError("'points' must be a list of 1D arrays!")
    if ndim != values.ndim:
        raise TypeError(f'Dimension mismatch! Grid is {ndim}D but values are {values.ndim}D!')
    if interp_points.ndim != 2:
        raise TypeError("'interp_points' should have shape (N, ndim).")
    if interp_points.shape[1] != ndim:
        raise TypeError(f'Dimension mismatch! Grid is {ndim}D but interpolation points are {interp_points.shape[1]}D!')

    fill = fill_value is not None
    n = interp_points.shape[0]
    his, los, bad = [], [], np.zeros(n, dtype=bool)
    for d in range(ndim):
        hi, b = _hi_index(points[d], interp_points[:, d], fill)
        his.append(hi)
        los.append(hi - 1)
        bad |= b

    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        if ndim == 1:
            x = interp_points[:, 0]
            x0, x1 = points[0][los[0]], points[0][his[0]]
            y0, y1 = values[los[0]], values[his[0]]
            slope = (y1 - y0) / (x1 - x0)
            out = y0 + slope * (x - x0)
        elif ndim == 2:
            x, y = interp_points[:, 0], interp_points[:, 1]
            x0, x1 = points[0][los[0]], points[0][his[0]]
            y0, y1 = points[1][los[1]], points[1][his[1]]
            z00 = values[los[0], los[1]]
            z01 = values[los[0], his[1]]
            z10 = values[his[0], los[1]]
            z11 = values[his[0], his[1]]
            dx, dy = x1 - x0, y1 - y0
            dx0, dx1, dy0, dy1 = x - x0, x1 - x, y - y0, y1 - y
            out = (dx1 * (z00 * dy1 + z01 * dy0) + dx0 * (z10 * dy1 + z11 * dy0)) / (dx * dy)
        elif ndim == 3:
            x, y, z = interp_points[:, 0], interp_points[:, 1], interp_points[:, 2]
            x0, x1 = points[0][los[0]], points[0][his[0]]
            y0, y1 = points[1][los[1]], points[1][his[1]]
            z0, z1 = points[2][los[2]], points[2][his[2]]
            w = lambda a, b, c: values[a, b, c]
            w000, w001 = w(los[0], los[1], los[2]), w(los[0], los[1], his[2])
            w010, w011 = w(los[0], his[1], los[2]), w(los[0], his[1], his[2])
            w100, w101 = w(his[0], los[1], los[2]), w(his[0], los[1], his[2])
            w110, w111 = w(his[0], his[1], los[2]), w(his[0], his[1], his[2])
            dx, dy, dz = x1 - x0, y1 - y0, z1 - z0
            dx0, dx1, dy0, dy1, dz0, dz1 = x - x0, x1 - x, y - y0, y1 - y, z - z0, z1 - z
            out = (
                dx1 * (dy1 * (dz1 * w000 + dz0 * w001) + dy0 * (dz1 * w010 + dz0 * w011))
                + dx0 * (dy1 * (dz1 * w100 + dz0 * w101) + dy0 * (dz1 * w110 + dz0 * w111))
            ) / (dx * dy * dz)
        else:
            total_volume = np.ones(n)
            lower, upper = [], []
            for d in range(ndim):
                x = interp_points[:, d]
                x0, x1 = points[d][los[d]], points[d][his[d]]
                total_volume = total_volume * (x1 - x0)
                lower.append(x - x0)
                upper.append(x1 - x)
            out = np.zeros(n)
            for j in range(1 << ndim):
                idx = tuple(his[d] if (j >> d) & 1 else los[d] for d in range(ndim))
                term = values[idx]
                for d in range(ndim):
                    term = term * (lower[d] if (j >> d) & 1 else upper[d])
                out = out + term
            out = out / total_volume
    if fill:
        out = np.where(bad, np.float64(fill_value), out)
    return out


def interpolate_along_axis(points, values, interp_points, axis=-1, fill_value=None, assume_sorted=False, max_threads=8):
    """RAiDER.interpolate.interpolate_along_axis (module.cpp:296-493 -> interpolate.cpp:260-332)."""
    points = np.asarray(points, dtype=np.float64, order='C')
    values = np.asarray(values, dtype=np.float64, order='C')
    interp_points = np.asarray(interp_points, dtype=np.float64, order='C')
    if values.ndim == 0 or interp_points.ndim == 0:
        raise TypeError('Only arrays are supported, not scalar values!')
    if points.ndim != values.ndim or points.ndim != interp_points.ndim:
        raise TypeError("'points', 'values' and 'interp_points' must all have the same number of dimensions!")
    ndim = points.ndim
    if points.shape != values.shape:
        raise TypeError("'points' and 'values' must have the same shape!")
    if axis < 0:
        axis += ndim
    if axis >= ndim or axis < 0:
        raise TypeError("'axis' out of range!")
    elif axis == 0 and max_threads > 1:
        raise RuntimeError('Cannot interpolate along axis 0 with multiple threads!')
    for i in range(ndim):
        if i != axis and interp_points.shape[i] != points.shape[i]:
            raise TypeError(
                f"Dimension mismatch at axis {i}! 'points' is {points.shape[i]} but interp_points is {interp_points.shape[i]}!"
            )

    fill = fill_value is not None
    g = np.moveaxis(points, axis, -1).reshape(-1, points.shape[axis])
    v = np.moveaxis(values, axis, -1).reshape(-1, points.shape[axis])
    q = np.moveaxis(interp_points, axis, -1)
    qshape = q.shape
    q = q.reshape(-1, interp_points.shape[axis])
    out = np.empty(q.shape)
    n = g.shape[1]
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        for c in range(g.shape[0]):
            hi = np.searchsorted(g[c], q[c], side='right') if _is_sorted(g[c]) else np.array([bisect_left(g[c], x) for x in q[c]])
            bad = (hi < 1) | (hi > n - 1)
            hi = np.clip(hi, 1, n - 1)
            lo = hi - 1
            x0, x1, y0, y1 = g[c][lo], g[c][hi], v[c][lo], v[c][hi]
            slope = (y1 - y0) / (x1 - x0)
            res = y0 + slope * (q[c] - x0)
            if fill:
                res = np.where(bad, np.float64(fill_value), res)
            out[c] = res
    return np.ascontiguousarray(np.moveaxis(out.reshape(qshape), -1, axis))


def _is_sorted(a):
    return bool(np.all(a[1:] >= a[:-1])) and not np.isnan(a).any()


def make_npts(max_len, step):
    """makePoints.pyx:130-134 *as Cython compiles it*: with C doubles ``a // b`` is ``floor(a / b)`` (not CPython's
    divmod-based float floor division) while ``a % b`` keeps Python's sign convention -- so (1.0, 0.1) gives 11, where
    pure Python would say 10.  Pinned against the compiled reference in tests/golden/makepoints.npz['counts']."""
    max_len = float(max_len)
    step = float(step)
    n = int(np.floor(max_len / step))
    if max_len % step != 0:
        n += 1
    return n


def makePoints(max_len, Rays_SP, Rays_SLV, stepSize):
    """makePoints{0,1,2,3}D (makePoints.pyx:15-148): ray[..., c, k] = SP[..., c] + basespace[k] * SLV[..., c]."""
    sp = np.asarray(Rays_SP)
    slv = np.asarray(Rays_SLV)
    if sp.dtype != np.float64 or slv.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'double'")
    npts = make_npts(max_len, stepSize)
    basespace = np.arange(0, float(max_len) + float(stepSize), float(stepSize))
    return sp[..., :, None] + basespace[:npts] * slv[..., :, None]


def scipy_find_interval(grid, x):
    """Interval rule of scipy RGI (find_interval_ascending, extrapolate=True; Appendix A of SURVEY.md).

    Returns (index, norm_distance, out_of_bounds): ``grid[i] <= x < grid[i+1]``, last node inclusive.
    """
    grid = np.asarray(grid, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = grid.shape[0]
    i = np.clip(np.searchsorted(grid, x, side='right') - 1, 0, n - 2)
    with np.errstate(invalid='ignore'):
        t = (x - grid[i]) / (grid[i + 1] - grid[i])
        oob = (x < grid[0]) | (x > grid[-1])
    return i, t, oob
