"""Random-geometry stress of the production integrator against the PROJ-form one: latitude, longitude, incidence, heading, ground
height, number of levels, segment length, cube spacing and posting drawn at random; per case the max |difference| of both delay
maps, the NaN pattern and the step counts must agree.  (Not a timing run.)

    python profiles/stress_forms.py [n_cases] [seed]
"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402,F401

from raider_b200 import synthetic as syn  # noqa: E402
from raider_b200.delay import _build_cube_ray  # noqa: E402
from raider_b200.delayFcns import getInterpolators  # noqa: E402
from raider_b200.losreader import Raytracing  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
worst = 0.0
for case in range(n_cases):
    lat0, lon0 = rng.uniform(-75, 75), rng.uniform(-179, 179)
    inc, head = rng.uniform(0.0, 58.0), rng.uniform(-180, 180)
    nz = int(rng.choice([24, 37, 50, 72]))
    table = rng.choice([None, None, 'ml145', 'hrrr57'])
    seg = float(rng.choice([150.0, 225.0, 400.0, 1000.0]))
    spacing = float(rng.choice([0.0625, 0.125, 0.25, 0.5]))
    posting = float(rng.choice([0.0005, 0.002, 0.01]))
    ny, nx = int(rng.choice([40, 64])), int(rng.choice([48, 61]))
    ht = float(rng.choice([0.0, 0.0, 350.0, 1800.0]))
    xp, yp = syn.raster(lat0, lon0, ny, nx, posting)
    xs, ys = syn.cube_axes_around(xp, yp, spacing_deg=spacing, pad_deg=max(2.5, 6 * spacing))
    zs = syn.z_levels_table(table) if table else syn.z_levels(nz)
    cube = syn.make_cube(ys, xs, zs, seed=int(rng.integers(1 << 30)), totals=False)
    zref = float(zs[-1] - 1.0)
    res = {}
    for mode in ('poly', 'general'):
        os.environ['RDR_K3_MODE'] = mode
        ifs = getInterpolators(cube)
        out = _build_cube_ray(xp, yp, np.array([ht]), Raytracing(incidence=inc, heading=head), 4326, 4326, list(ifs), MAX_SEGMENT_LENGTH=seg, MAX_TROPO_HEIGHT=zref)
        res[mode] = (out, ifs[0].cube.last_info[0])
    (a, ia), (b, ib) = res['poly'], res['general']
    same_nan = np.array_equal(np.isnan(a[0]), np.isnan(b[0])) and np.array_equal(np.isnan(a[1]), np.isnan(b[1]))
    d = max(np.nanmax(np.abs(a[0] - b[0])) if np.isfinite(b[0]).any() else 0.0, np.nanmax(np.abs(a[1] - b[1])) if np.isfinite(b[1]).any() else 0.0)
    worst = max(worst, d)
    flag = '' if (same_nan and d < 1e-9 and np.array_equal(ia.nparts, ib.nparts)) else '   <-- CHECK'
    print(f'case {case:2d}: lat {lat0:6.1f} inc {inc:4.1f} head {head:6.1f} ht {ht:6.0f} levels {zs.size:3d} seg {seg:5.0f} cells {spacing} posting {posting}: '
          f'samples/ray {ia.samples_per_ray:4d} nan {int(np.isnan(b[0]).sum()):5d} max|d| {d:.2e} nan-pattern {same_nan}{flag}', flush=True)
print('worst max|d|', worst)
