"""Synthetic weather-model cubes and query rasters -- pure NumPy, no oracle, no GPU.

These are the inputs bench.py and the parity tests run on (there is no network, hence no real
ERA5/HRRR/GMAO product).  A *cube* is the in-memory form of the processed weather-model file the
reference's ``getInterpolators`` reads (tools/RAiDER/delayFcns.py:31-41; written by
tools/RAiDER/models/weatherModel.py:659-724): ``x, y, z`` coordinate vectors and float32 fields
``wet, hydro, wet_total, hydro_total`` with shape (z, y, x).
"""
from __future__ import annotations

import numpy as np

_P0_HPA, _T0_K = 1013.25, 288.15


def z_levels(nz: int) -> np.ndarray:
    """Nominal z nodes: zs[k] = -500 + 48500 (k/(NZ-1))^2, strictly increasing f64 (NZ = 37/50/72)."""
    k = np.arange(nz, dtype=np.float64)
    return -500.0 + 48500.0 * (k / (nz - 1)) ** 2


def z_levels_table(kind: str) -> np.ndarray:
    """Height tables shaped like the reference's processed cubes ("raw levels vs z nodes").

    'ml145': 145 nodes, -500 .. ~80.3 km, thickness ~20 m near the ground, ~300 m in the upper troposphere,
             km-scale above 40 km (the shape of the ERA5/GMAO model-level table, models/model_levels.py:12).
    'hrrr57': 57 nodes, -500 .. ~26.2 km (shape of models/model_levels.py:517).
    The numbers are generated here (smooth stretching), not copied from the reference tables.
    """
    if kind == 'ml145':
        neg = np.array([-500.0, -300.0, -200.0, -100.0, -50.0, -20.0])
        n_pos, top = 145 - neg.size, 80301.65
        i = np.arange(n_pos - 1, dtype=np.float64)
        thick = 20.0 + 280.0 * (1.0 - np.exp(-i / 22.0))
        thick *= np.where(i > 88, np.exp((i - 88) / 15.5), 1.0)
    elif kind == 'hrrr57':
        neg = np.array([-500.0, -200.0, -100.0, -50.0, -20.0, -10.0])
        n_pos, top = 57 - neg.size, 26158.0385
        i = np.arange(n_pos - 1, dtype=np.float64)
        thick = 20.0 + 1100.0 * (1.0 - np.exp(-i / 14.0))
    else:
        raise ValueError(f'unknown level table {kind!r}')
    pos = 10.0 + np.concatenate([[0.0], np.cumsum(thick)])
    pos = 10.0 + (pos - 10.0) * ((top - 10.0) / (pos[-1] - 10.0))
    zs = np.concatenate([neg, pos])
    assert zs.size == neg.size + n_pos and np.all(np.diff(zs) > 0)
    return zs


def cumulative_total(field_zyx: np.ndarray, zs: np.ndarray) -> np.ndarray:
    """``_getZTD`` (weatherModel.py:389-403): total[k] = 1e-6 * trapz(field[k:], zs[k:]) as a reversed cumulative sum."""
    f = np.asarray(field_zyx, dtype=np.float64)
    dz = np.diff(np.asarray(zs, dtype=np.float64))[:, None, None]
    seg = 0.5 * (f[1:] + f[:-1]) * dz
    out = np.zeros(f.shape)
    out[:-1] = np.cumsum(seg[::-1], axis=0)[::-1]
    return 1e-6 * out


def make_cube(ys, xs, zs, seed: int = 20200130, noise: float = 0.5, lat_of=None, lon_of=None, totals: bool = True) -> dict:
    """Analytic + noise refractivity cube.

    hydro = 77.6 P0/T0 exp(-h/8000) (1 + 0.02 sin(lat));  wet = 60 exp(-h/2000) (1 + 0.3 cos(lon pi/90));
    both + noise * N(0,1), rng(seed), cast to float32.  For projected grids pass ``lat_of``/``lon_of``
    (ny, nx) arrays giving the geographic position of the nodes.
    """
    ys = np.asarray(ys, dtype=np.float64)
    xs = np.asarray(xs, dtype=np.float64)
    zs = np.asarray(zs, dtype=np.float64)
    lat = np.broadcast_to(ys[:, None], (ys.size, xs.size)) if lat_of is None else np.asarray(lat_of)
    lon = np.broadcast_to(xs[None, :], (ys.size, xs.size)) if lon_of is None else np.asarray(lon_of)
    h = zs[:, None, None]
    hydro = 77.6 * _P0_HPA / _T0_K * np.exp(-h / 8000.0) * (1.0 + 0.02 * np.sin(np.radians(lat)))[None]
    wet = 60.0 * np.exp(-h / 2000.0) * (1.0 + 0.3 * np.cos(lon * np.pi / 90.0))[None]
    rng = np.random.default_rng(seed)
    wet = wet + noise * rng.standard_normal(wet.shape)
    hydro = hydro + noise * rng.standard_normal(hydro.shape)
    cube = {'x': xs, 'y': ys, 'z': zs, 'wet': wet.astype(np.float32), 'hydro': hydro.astype(np.float32)}
    if totals:
        cube['wet_total'] = cumulative_total(cube['wet'], zs).astype(np.float32)
        cube['hydro_total'] = cumulative_total(cube['hydro'], zs).astype(np.float32)
    return cube


def constant_cube(ys, xs, zs, wet_value: float, hydro_value: float) -> dict:
    """Constant-refractivity cube: the known-answer set-up of test/test_synthetic.py:217-274 (delay = N 1e-6 sum L)."""
    shape = (np.size(zs), np.size(ys), np.size(xs))
    return {
        'x': np.asarray(xs, dtype=np.float64), 'y': np.asarray(ys, dtype=np.float64), 'z': np.asarray(zs, dtype=np.float64),
        'wet': np.full(shape, wet_value, dtype=np.float32), 'hydro': np.full(shape, hydro_value, dtype=np.float32),
    }


def blend_cubes(c0: dict, c1: dict, w0: float, w1: float) -> dict:
    """Temporal interpolation as the reference does it: cube = sum_i w_i cube_i per field (cli/raider.py:817-819)."""
    out = {'x': c0['x'], 'y': c0['y'], 'z': c0['z']}
    for k in ('wet', 'hydro', 'wet_total', 'hydro_total'):
        if k in c0 and k in c1:
            out[k] = (w0 * c0[k].astype(np.float64) + w1 * c1[k].astype(np.float64)).astype(np.float32)
    return out


def raster(center_lat: float, center_lon: float, ny: int, nx: int, posting_deg: float):
    """Query grid as the reference AOI builds it (llreader.py:190-191): xpts ascending, ypts descending."""
    x0 = center_lon - 0.5 * (nx - 1) * posting_deg
    y0 = center_lat + 0.5 * (ny - 1) * posting_deg
    xpts = x0 + posting_deg * np.arange(nx, dtype=np.float64)
    ypts = y0 - posting_deg * np.arange(ny, dtype=np.float64)
    return xpts, ypts


def cube_axes_around(xpts, ypts, spacing_deg: float = 0.25, pad_deg: float = 2.0):
    """0.25-degree cube axes padded ``pad_deg`` beyond the raster on all sides, snapped to the spacing."""
    lo_x = np.floor((np.min(xpts) - pad_deg) / spacing_deg) * spacing_deg
    hi_x = np.ceil((np.max(xpts) + pad_deg) / spacing_deg) * spacing_deg
    lo_y = np.floor((np.min(ypts) - pad_deg) / spacing_deg) * spacing_deg
    hi_y = np.ceil((np.max(ypts) + pad_deg) / spacing_deg) * spacing_deg
    xs = lo_x + spacing_deg * np.arange(int(round((hi_x - lo_x) / spacing_deg)) + 1, dtype=np.float64)
    ys = lo_y + spacing_deg * np.arange(int(round((hi_y - lo_y) / spacing_deg)) + 1, dtype=np.float64)
    return xs, ys


# BASELINE.json configs.  C2 is the configuration the headline metric is quoted on.
def config_c1():
    """C1: zenith, 100x100 query grid over an 11x15x37 cube at 0.25 deg (test/scenario_1 shape)."""
    ys = 15.75 + 0.25 * np.arange(11)
    xs = -103.25 + 0.25 * np.arange(15)
    cube = make_cube(ys, xs, z_levels(37))
    xpts = np.linspace(xs[0] + 0.1, xs[-1] - 0.1, 100)
    ypts = np.linspace(ys[-1] - 0.1, ys[0] + 0.1, 100)
    return {'cube': cube, 'xpts': xpts, 'ypts': ypts, 'zpts': np.array([0.0, 50.0, 100.0, 500.0, 1000.0])}


def config_c2(n: int = 2000, nz: int = 37, table: str | None = None, max_segment_length: float | None = None):
    """C2: slant delay, fixed 30 deg incidence (heading -168), n x n raster at 0.001 deg centred (34,-118), ht = 0."""
    xpts, ypts = raster(34.0, -118.0, n, n, 0.001)
    # the cube is always laid out for the full 2000 x 2000 footprint so sub-rasters see the same grid
    fx, fy = raster(34.0, -118.0, 2000, 2000, 0.001)
    xs, ys = cube_axes_around(fx, fy)
    zs = z_levels(nz) if table is None else z_levels_table(table)
    cube = make_cube(ys, xs, zs)
    if max_segment_length is None:
        max_segment_length = 225.0 if table is None else 1000.0
    return {
        'cube': cube, 'xpts': xpts, 'ypts': ypts, 'zpts': np.array([0.0]), 'incidence': 30.0, 'heading': -168.0,
        'zref': float(zs[-1] - 1.0), 'max_segment_length': float(max_segment_length),
    }


def config_c4(n: int = 10000, seed: int = 7):
    """C4: GNSS point mode -- ``n`` stations uniform in a 4 x 4 degree box, heights U(0, 3000) m, per-station incidence
    U(5, 75) deg and heading U(0, 360) deg, two weather epochs (seeds 20200130 / 20200131) blended 0.5 / 0.5."""
    rng = np.random.default_rng(seed)
    lat = rng.uniform(32.0, 36.0, n)
    lon = rng.uniform(-120.0, -116.0, n)
    hgt = rng.uniform(0.0, 3000.0, n)
    inc = rng.uniform(5.0, 75.0, n)
    head = rng.uniform(0.0, 360.0, n)
    xs, ys = cube_axes_around(np.array([-120.0, -116.0]), np.array([32.0, 36.0]), pad_deg=3.0)
    zs = z_levels(37)
    c0 = make_cube(ys, xs, zs, seed=20200130, totals=False)
    c1 = make_cube(ys, xs, zs, seed=20200131, totals=False)
    # zref well below the model top: the fixed-point layer tops of 75-degree rays overshoot by metres (losreader.py:713-716)
    return {'cube0': c0, 'cube1': c1, 'weights': (0.5, 0.5), 'lat': lat, 'lon': lon, 'hgt': hgt, 'incidence': inc, 'heading': head,
            'zref': 15000.0, 'max_segment_length': 1000.0}


def circular_orbit(lat0_deg: float, lon0_deg: float, heading_deg: float = -168.0, h_sat: float = 700000.0, omega_deg_s: float = 0.06,
                   look_angle_deg: float = 33.0, n_sv: int = 41, dt: float = 10.0):
    """State vectors of a circular orbit whose right-looking zero-Doppler footprint passes over (lat0, lon0) at mid time.

    The construction of test/fake_raytracing:73-117 (circle of radius a + h_sat, constant angular rate) generalised to an
    inclined great circle: the sub-satellite track heads along ``heading_deg`` (clockwise from north) and is offset to the
    left of the target so that the target is seen at roughly ``look_angle_deg`` off nadir.  Returns rows (t, x, y, z, vx, vy, vz).
    """
    a = 6378137.0
    r = a + h_sat
    lat0, lon0, hd = np.radians(lat0_deg), np.radians(lon0_deg), np.radians(heading_deg)
    up = np.array([np.cos(lat0) * np.cos(lon0), np.cos(lat0) * np.sin(lon0), np.sin(lat0)])
    east = np.array([-np.sin(lon0), np.cos(lon0), 0.0])
    north = np.cross(up, east)
    along = np.sin(hd) * east + np.cos(hd) * north           # flight direction at the target's abeam point
    right = np.cross(along, up)                              # right of the flight direction (pointing to the target side)
    # sub-satellite point: ground range to the left of the target
    gr = np.radians(np.degrees(np.arcsin(r / a * np.sin(np.radians(look_angle_deg)))) - look_angle_deg)
    nadir = np.cos(gr) * up - np.sin(gr) * right
    om = np.radians(omega_deg_s)
    t = dt * np.arange(n_sv)
    ang = om * (t - t[n_sv // 2])
    pos = r * (np.cos(ang)[:, None] * nadir[None] + np.sin(ang)[:, None] * along[None])
    vel = r * om * (-np.sin(ang)[:, None] * nadir[None] + np.cos(ang)[:, None] * along[None])
    return np.concatenate([t[:, None], pos, vel], axis=1)


def raster_xy(center_lat: float, center_lon: float, ny: int, nx: int, posting_y: float, posting_x: float):
    """Like :func:`raster` with different postings along y and x."""
    xpts = center_lon - 0.5 * (nx - 1) * posting_x + posting_x * np.arange(nx, dtype=np.float64)
    ypts = center_lat + 0.5 * (ny - 1) * posting_y - posting_y * np.arange(ny, dtype=np.float64)
    return xpts, ypts


def config_c3(ny: int = 8000, nx: int = 10000, nz: int = 50, table: str | None = None):
    """C3: per-pixel LOS from orbit state vectors over a Sentinel-1 IW-swath-sized raster (2.5 x 1.8 degrees, 8e7 points at
    full size), HRRR-like cube: 3 km spherical-LCC grid (models/hrrr.py:255-260) padded 15 cells, NZ = 50 (or 'hrrr57')."""
    from .crs import LambertConformalSphere
    lcc = LambertConformalSphere()
    xpts, ypts = raster_xy(36.0, -98.0, ny, nx, 1.8 / ny, 2.5 / nx)
    cx, cy = lcc.from_ll(np.array([xpts[0], xpts[-1], xpts[0], xpts[-1], -98.0, -98.0]),
                         np.array([ypts[0], ypts[0], ypts[-1], ypts[-1], ypts[0], ypts[-1]]))
    pad = 15 * 3000.0
    x0, x1 = np.floor((cx.min() - pad) / 3000.0) * 3000.0, np.ceil((cx.max() + pad) / 3000.0) * 3000.0
    y0, y1 = np.floor((cy.min() - pad) / 3000.0) * 3000.0, np.ceil((cy.max() + pad) / 3000.0) * 3000.0
    xs = x0 + 3000.0 * np.arange(int(round((x1 - x0) / 3000.0)) + 1)
    ys = y0 + 3000.0 * np.arange(int(round((y1 - y0) / 3000.0)) + 1)
    X, Y = np.meshgrid(xs, ys)
    lon_n, lat_n, _ = lcc.to_llh(X, Y, np.zeros_like(X))
    zs = z_levels(nz) if table is None else z_levels_table(table)
    cube = make_cube(ys, xs, zs, lat_of=lat_n, lon_of=lon_n, seed=20200130, totals=False)
    cube['crs'] = lcc
    return {'cube': cube, 'crs': lcc, 'xpts': xpts, 'ypts': ypts, 'zpts': np.array([0.0]), 'orbit_rows': circular_orbit(36.0, -98.0),
            'zref': float(zs[-1] - 1.0), 'max_segment_length': 1000.0}


def config_c5(ny: int = 12000, nx: int = 20000, nz: int = 72, table: str | None = None):
    """C5: NISAR-scale raster (2.4e8 pixels at full size, 0.0002 deg posting), GMAO-like cube 0.25 x 0.3125 degrees
    (models/gmao.py:47-48), NZ = 72 (or 'ml145'), fixed 30 deg incidence as C2; row-block sharded over the GPUs."""
    xpts, ypts = raster_xy(34.0, -118.0, ny, nx, 2.4 / ny, 4.0 / nx)
    lo_x, hi_x = np.floor((xpts[0] - 2.0) / 0.3125) * 0.3125, np.ceil((xpts[-1] + 2.0) / 0.3125) * 0.3125
    lo_y, hi_y = np.floor((ypts[-1] - 2.0) / 0.25) * 0.25, np.ceil((ypts[0] + 2.0) / 0.25) * 0.25
    xs = lo_x + 0.3125 * np.arange(int(round((hi_x - lo_x) / 0.3125)) + 1)
    ys = lo_y + 0.25 * np.arange(int(round((hi_y - lo_y) / 0.25)) + 1)
    zs = z_levels(nz) if table is None else z_levels_table(table)
    cube = make_cube(ys, xs, zs, seed=20200130, totals=False)
    return {'cube': cube, 'xpts': xpts, 'ypts': ypts, 'zpts': np.array([0.0]), 'incidence': 30.0, 'heading': -168.0,
            'zref': float(zs[-1] - 1.0), 'max_segment_length': 1000.0}
