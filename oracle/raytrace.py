"""CPU restatement of the RAiDER ray-tracing delay path -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows, loop for loop, the reference (paths relative to /root/reference):

* :func:`getTopOfAtmosphere`  tools/RAiDER/losreader.py:706-733
* :func:`build_ray`           tools/RAiDER/losreader.py:772-835
* :func:`build_cube_ray`      tools/RAiDER/delay.py:219-326   (``_build_cube_ray``)
* :func:`build_cube`          tools/RAiDER/delay.py:196-216   (``_build_cube``)
* :func:`get_interpolators`   tools/RAiDER/delayFcns.py:23-58 (scipy RGI, fill_value=nan,
  bounds_error=False, fp32 values in a transposed (y,x,z) view)

PIN STATUS: pinned BIT FOR BIT to the reference's own Python.  ``oracle/refpy.py`` imports RAiDER.delay / losreader /
delayFcns / utilFcns unmodified from /root/reference (stand-ins only for pyproj / xarray / rasterio / shapely) and
``tests/test_oracle_vs_reference_py.py`` asserts ``np.array_equal`` between every function below and the reference's
on the golden geometries, edge rules and error rules; ``tests/golden/raytrace.npz`` is written from the reference
functions' outputs.  What stays unpinned is PROJ's own rounding (the pyproj stand-in computes with oracle.geodesy).

Differences from the reference, all forced by what is importable offline:

* PROJ transforms are replaced by :mod:`oracle.geodesy` (PROJ's published ``cart``/``lcc`` algorithms).
* ``los.getLookVectors`` (isce3 per-pixel geo2rdr, losreader.py:219-255) is replaced by a duck-typed
  LOS provider: any object with ``getLookVectors(ht, llh, xyz, yy) -> (ny, nx, 3)``.
  :class:`FixedIncidenceLOS` (losreader.py:374-396 + utilFcns.py:91-121) and :class:`ZenithLOS`
  (losreader.py:302-316) are provided; arbitrary arrays via :class:`ArrayLOS`.
* ``layer_maxlen`` may be injected so that a sub-raster reproduces the full raster's global
  ``nParts`` (delay.py:283) -- used only by the bounded CPU-baseline sample in bench.py.
"""
from __future__ import annotations

import numpy as np
from scipy.interpolate import RegularGridInterpolator as Interpolator

from . import geodesy

_ZREF = np.float64(26000)  # tools/RAiDER/constants.py:12


# ------------------------------------------------------------------------------------------
# LOS providers (duck type used at delay.py:270)
# ------------------------------------------------------------------------------------------
class ZenithLOS:
    """Zenith look vectors in ECEF, losreader.py:302-316."""

    def getLookVectors(self, ht, llh, xyz, yy):
        return geodesy.getZenithLookVecs(llh[1], llh[0], llh[2])


class FixedIncidenceLOS:
    """Constant incidence/heading: inc_hd_to_enu (losreader.py:374-396) -> enu2ecef (utilFcns.py:91-121) per pixel."""

    def __init__(self, incidence_deg: float, heading_deg: float) -> None:
        self.incidence_deg, self.heading_deg = float(incidence_deg), float(heading_deg)
        self.enu = geodesy.inc_hd_to_enu(np.float64(incidence_deg), np.float64(heading_deg))

    def getLookVectors(self, ht, llh, xyz, yy):
        e, n, u = self.enu
        return geodesy.enu2ecef(e, n, u, llh[1], llh[0], llh[2])


class ArrayLOS:
    """Explicit (ny, nx, 3) ECEF unit vectors (what Raytracing.getLookVectors returns, losreader.py:219-255)."""

    def __init__(self, vecs) -> None:
        self.vecs = np.asarray(vecs, dtype=np.float64)

    def getLookVectors(self, ht, llh, xyz, yy):
        return self.vecs


# ------------------------------------------------------------------------------------------
# model CRS (delay.py:253 ``ecef_to_model``)
# ------------------------------------------------------------------------------------------
class GeographicCRS:
    """EPSG:4326 model grid: ecef -> (lon, lat, h)."""
    kind = 0

    def ecef_to_model(self, x, y, z):
        return geodesy.ecef2lla(x, y, z)

    def model_to_llh(self, xx, yy, hh):
        return [xx, yy, hh]

    def params(self):
        return np.zeros(7)

    def __eq__(self, other):
        return isinstance(other, GeographicCRS)


class LambertCRS(GeographicCRS):
    """HRRR-style spherical LCC model grid: ecef -> (x_m, y_m, h)."""
    kind = 1

    def __init__(self, **kw) -> None:
        self.lcc = geodesy.LambertConformalSphere(**kw)

    def ecef_to_model(self, x, y, z):
        lon, lat, h = geodesy.ecef2lla(x, y, z)
        X, Y = self.lcc.forward(lon, lat)
        return X, Y, h

    def model_to_llh(self, xx, yy, hh):
        lon, lat = self.lcc.inverse(xx, yy)
        return [lon, lat, hh]

    def params(self):
        return self.lcc.params()

    def __eq__(self, other):
        return isinstance(other, LambertCRS) and np.array_equal(self.params(), other.params())


# ------------------------------------------------------------------------------------------
def get_interpolators(cube: dict, kind: str = 'pointwise'):
    """delayFcns.py:23-58 on an in-memory cube {x, y, z, wet, hydro[, wet_total, hydro_total]} with (z, y, x) fields."""
    xs_wm = np.array(cube['x'])
    ys_wm = np.array(cube['y'])
    zs_wm = np.array(cube['z'])
    wet = cube['wet_total' if kind == 'total' else 'wet']
    hydro = cube['hydro_total' if kind == 'total' else 'hydro']
    wet = np.array(wet).transpose(1, 2, 0)
    hydro = np.array(hydro).transpose(1, 2, 0)
    ifWet = Interpolator((ys_wm, xs_wm, zs_wm), wet, fill_value=np.nan, bounds_error=False)
    ifHydro = Interpolator((ys_wm, xs_wm, zs_wm), hydro, fill_value=np.nan, bounds_error=False)
    return ifWet, ifHydro


def getTopOfAtmosphere(xyz, look_vecs, toaheight, factor=None):
    """losreader.py:706-733 -- Newton-Raphson along the ray to geodetic height ``toaheight``."""
    if factor is not None:
        maxIter = 3
    else:
        maxIter = 10
        factor = 1.0

    pos = xyz + toaheight * look_vecs

    for _ in range(maxIter):
        pos_llh = geodesy.ecef2lla(pos[..., 0], pos[..., 1], pos[..., 2])
        pos = pos + look_vecs * ((toaheight - pos_llh[2]) / factor)[..., None]

    return pos


def build_ray(model_zs, ht, xyz, LOS, MAX_TROPO_HEIGHT=_ZREF):
    """losreader.py:772-835 -- per model layer: segment end points and lengths (bottom up)."""
    low_xyz = None
    high_xyz = None
    cos_factor = None

    ray_lengths, low_xyzs, high_xyzs = [], [], []
    for zz in range(model_zs.size - 1):
        low_ht = model_zs[zz]
        high_ht = model_zs[zz + 1]

        if high_ht == model_zs[-1]:
            high_ht -= 0.01

        if (high_ht < ht) or (low_ht >= MAX_TROPO_HEIGHT):
            continue

        if low_ht < ht:
            low_ht = ht

        if high_ht > MAX_TROPO_HEIGHT:
            high_ht = MAX_TROPO_HEIGHT

        if np.abs(high_ht - low_ht) < 1.0:
            continue

        if high_xyz is not None:
            low_xyz = high_xyz
        else:
            low_xyz = getTopOfAtmosphere(xyz, LOS, low_ht, factor=cos_factor)

        high_xyz = getTopOfAtmosphere(xyz, LOS, high_ht, factor=cos_factor)

        ray_length = np.linalg.norm(high_xyz - low_xyz, axis=-1)

        if cos_factor is None:
            cos_factor = (high_ht - low_ht) / ray_length

        ray_lengths.append(ray_length)
        low_xyzs.append(low_xyz)
        high_xyzs.append(high_xyz)

    if not ray_lengths:
        return None, None, None
    else:
        return np.stack(ray_lengths), np.stack(low_xyzs), np.stack(high_xyzs)


def layer_plan(model_zs, ht, MAX_TROPO_HEIGHT=_ZREF):
    """The scalar (pixel-independent) layer decisions of build_ray, losreader.py:785-809, as a list of (low_ht, high_ht)."""
    model_zs = np.asarray(model_zs)
    out = []
    for zz in range(model_zs.size - 1):
        low_ht = model_zs[zz]
        high_ht = model_zs[zz + 1]
        if high_ht == model_zs[-1]:
            high_ht -= 0.01
        if (high_ht < ht) or (low_ht >= MAX_TROPO_HEIGHT):
            continue
        if low_ht < ht:
            low_ht = ht
        if high_ht > MAX_TROPO_HEIGHT:
            high_ht = MAX_TROPO_HEIGHT
        if np.abs(high_ht - low_ht) < 1.0:
            continue
        out.append((float(low_ht), float(high_ht)))
    return out


def n_parts(ray_lengths, MAX_SEGMENT_LENGTH=1000.0, layer_maxlen=None):
    """delay.py:283."""
    mx = ray_lengths.max((1, 2)) if layer_maxlen is None else np.asarray(layer_maxlen)
    return np.ceil(mx / MAX_SEGMENT_LENGTH).astype(int) + 1


def build_cube_ray(
    xpts,
    ypts,
    zpts,
    los,
    model_crs,
    pts_crs,
    interpolators,
    outputArrs=None,
    MAX_SEGMENT_LENGTH=1000.0,
    MAX_TROPO_HEIGHT=_ZREF,
    layer_maxlen=None,
    stats=None,
):
    """delay.py:219-326 (``_build_cube_ray``).

    ``layer_maxlen`` (optional, list over zpts of per-layer maxima) overrides the raster max at
    delay.py:283; ``stats`` (optional dict) receives nParts / sample counts per height.
    """
    model_zs = interpolators[0].grid[2]
    xx, yy = np.meshgrid(xpts, ypts)
    zpts = np.asarray(zpts)

    output_created_here = False
    if outputArrs is None:
        output_created_here = True
        outputArrs = [np.zeros((zpts.size, ypts.size, xpts.size)) for mm in range(len(interpolators))]

    geographic = GeographicCRS()

    for hh, ht in enumerate(zpts):
        outSubs = [x[hh, ...] for x in outputArrs]

        # Step 1: transform points to llh and xyz (delay.py:262-267)
        if pts_crs != geographic:
            llh = list(pts_crs.model_to_llh(xx, yy, np.full(yy.shape, ht)))
        else:
            llh = [xx, yy, np.full(yy.shape, ht)]
        xyz = np.stack(geodesy.lla2ecef(llh[1], llh[0], llh[2]), axis=-1)

        # Step 2 - LOS vectors (delay.py:270)
        LOS = los.getLookVectors(ht, llh, xyz, yy)

        # Step 3 - ray segments per model layer (delay.py:273)
        ray_lengths, low_xyzs, high_xyzs = build_ray(model_zs, ht, xyz, LOS, MAX_TROPO_HEIGHT)

        if ray_lengths is None and ht == zpts[-1]:
            continue
        elif ray_lengths is None:
            # the reference evaluates np.isnan(None) here -> TypeError (latent bug, delay.py:279)
            raise TypeError('no model layer contributes at a height that is not the last output level')
        elif np.isnan(ray_lengths).all():
            raise ValueError('geo2rdr did not converge. Check orbit coverage')

        nParts = n_parts(ray_lengths, MAX_SEGMENT_LENGTH, None if layer_maxlen is None else layer_maxlen[hh])
        if stats is not None:
            stats.setdefault('nParts', []).append(nParts.copy())
            stats.setdefault('maxlen', []).append(ray_lengths.max((1, 2)))

        for zz, nparts in enumerate(nParts):
            fracs = np.linspace(0.0, 1.0, num=nparts)

            for findex, ff in enumerate(fracs):
                pts_xyz = low_xyzs[zz] + ff * (high_xyzs[zz] - low_xyzs[zz])

                pts = model_crs.ecef_to_model(pts_xyz[..., 0], pts_xyz[..., 1], pts_xyz[..., 2])
                pts = np.stack((pts[1], pts[0], pts[2]), axis=-1)

                if (pts[:, :, -1] < np.array(model_zs).min()).all():
                    pts[:, :, -1] = np.array(model_zs).min()

                if (pts[:, :, -1] > np.array(model_zs).max()).all():
                    pts[:, :, -1] = np.array(model_zs).max()

                wt = 0.5 if findex in [0, fracs.size - 1] else 1.0
                wt *= ray_lengths[zz] * 1.0e-6 / (nparts - 1.0)

                for mm, out in enumerate(outSubs):
                    val = interpolators[mm](pts)
                    out += wt * val

    if output_created_here:
        return outputArrs


def build_cube(xpts, ypts, zpts, model_crs, pts_crs, interpolators):
    """delay.py:196-216 (``_build_cube``): zenith / projected delays = one RGI call per field per height."""
    xx, yy = np.meshgrid(xpts, ypts)
    zpts = np.asarray(zpts)
    outputArrs = [np.zeros((zpts.size, ypts.size, xpts.size)) for mm in range(len(interpolators))]

    for ii, ht in enumerate(zpts):
        if model_crs != pts_crs:
            # transformPoints(yy, xx, ht, pts_crs, model_crs) -> (y, x, z) in the model system (delay.py:404-436)
            lon, lat, h = pts_crs.model_to_llh(xx, yy, np.full(yy.shape, ht))
            if model_crs.kind == 1:
                X, Y = model_crs.lcc.forward(lon, lat)
            else:
                X, Y = lon, lat
            pts = np.stack([Y, X, h], axis=-1)
        else:
            pts = np.stack([yy, xx, np.full(yy.shape, ht)], axis=-1)

        for mm, intp in enumerate(interpolators):
            outputArrs[mm][ii, ...] = intp(pts)

    return outputArrs


def cumulative_ztd(field_zyx, zs):
    """weatherModel.py:389-403 (``_getZTD``) on a (z, y, x) field: total[k] = 1e-6 * trapz(field[k:], zs[k:])."""
    f = np.asarray(field_zyx, dtype=np.float64)
    zs = np.asarray(zs, dtype=np.float64)
    out = np.zeros(f.shape)
    for level in range(f.shape[0]):
        out[level] = 1e-6 * np.trapezoid(f[level:], x=zs[level:], axis=0)
    return out
