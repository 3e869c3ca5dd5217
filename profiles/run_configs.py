"""BASELINE.json configs C1..C5 at full size on one GPU through the public API (host buffers in, host arrays out):
rays/s, samples per ray, NaN count, checksum.  C2 is the bench.py workload; the others are parity-test shapes (tests/ run them
at oracle-feasible sizes) and are measured here once per round for the record.

    python profiles/run_configs.py [c1 c2 c3 c4 c5] [--scale 1.0]     # --scale shrinks the raster side lengths
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402,F401  (CUDA context + pinned allocator warm-up)

from raider_b200 import synthetic as syn  # noqa: E402
from raider_b200.delay import _build_cube, _build_cube_ray, slant_delay_points  # noqa: E402
from raider_b200.delayFcns import getInterpolators  # noqa: E402
from raider_b200.losreader import Orbit, Raytracing  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith('--')]
scale = float(sys.argv[sys.argv.index('--scale') + 1]) if '--scale' in sys.argv else 1.0
which = args or ['c1', 'c2', 'c3', 'c4', 'c5']
out = {}


def timed(fn, reps=3):
    fn()  # warm-up: pools, pinned result arrays
    ts = []
    res = None
    for _ in range(reps):
        res = None  # steady state of a processing loop: the previous result is released, its page-locked block is reused
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return res, float(np.median(ts))


def report(name, n_rays, dt, res, extra):
    w, h = res
    rec = {'rays': int(n_rays), 'seconds': dt, 'rays_per_s': n_rays / dt, 'nan': int(np.isnan(w).sum()), 'checksum': float(np.nansum(w) + np.nansum(h))}
    rec.update(extra)
    out[name] = rec
    print(name, json.dumps(rec), flush=True)


if 'c1' in which:
    cfg = syn.config_c1()
    ifs = getInterpolators(cfg['cube'], 'total')
    res, dt = timed(lambda: _build_cube(cfg['xpts'], cfg['ypts'], cfg['zpts'], 4326, 4326, list(ifs)), reps=5)
    report('C1 zenith 100x100x5 heights, 11x15x37 cube', cfg['xpts'].size * cfg['ypts'].size * cfg['zpts'].size, dt, res, {'api': '_build_cube'})

if 'c2' in which:
    n = int(2000 * scale)
    cfg = syn.config_c2(n=n)
    los = Raytracing(incidence=30.0, heading=-168.0)

    def run():
        ifs = getInterpolators(cfg['cube'])
        r = _build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], los, 4326, 4326, list(ifs), MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                            MAX_TROPO_HEIGHT=cfg['zref'])
        run.info = ifs[0].cube.last_info[0]
        return r
    res, dt = timed(run)
    report(f'C2 slant {n}x{n}, 30 deg, NZ=37, 225 m', n * n, dt, res, {'samples_per_ray': run.info.samples_per_ray, 'api': 'getInterpolators + _build_cube_ray'})

if 'c3' in which:
    ny, nx = int(8000 * scale), int(10000 * scale)
    cfg = syn.config_c3(ny=ny, nx=nx)
    rows = cfg['orbit_rows']
    los = Raytracing(filename=Orbit(rows[:, 0], rows[:, 1:4], rows[:, 4:7]))

    def run():
        ifs = getInterpolators(cfg['cube'])
        r = _build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], los, cfg['crs'], 4326, list(ifs), MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                            MAX_TROPO_HEIGHT=cfg['zref'])
        run.info = ifs[0].cube.last_info[0]
        return r
    res, dt = timed(run, reps=2)
    report(f'C3 slant {ny}x{nx}, orbit LOS on device, HRRR-like LCC cube {cfg["cube"]["wet"].shape}', ny * nx, dt, res,
           {'samples_per_ray': run.info.samples_per_ray, 'tiles': run.info.tiles, 'api': 'getInterpolators + _build_cube_ray'})

if 'c4' in which:
    cfg = syn.config_c4(n=10000)
    los = Raytracing(incidence=cfg['incidence'], heading=cfg['heading'])

    def run():
        return slant_delay_points(cfg['cube0'], cfg['lat'], cfg['lon'], cfg['hgt'], los, zref=cfg['zref'], MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                                  second_epoch=cfg['cube1'], weights=cfg['weights'])
    res, dt = timed(run, reps=5)
    report('C4 10 000 stations, per-station height/LOS, two-epoch blend', 10000, dt, res, {'api': 'slant_delay_points'})

if 'c5' in which:
    ny, nx = int(12000 * scale), int(20000 * scale)
    cfg = syn.config_c5(ny=ny, nx=nx)
    los = Raytracing(incidence=30.0, heading=-168.0)

    def run():
        ifs = getInterpolators(cfg['cube'])
        r = _build_cube_ray(cfg['xpts'], cfg['ypts'], cfg['zpts'], los, 4326, 4326, list(ifs), MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                            MAX_TROPO_HEIGHT=cfg['zref'])
        run.info = ifs[0].cube.last_info[0]
        return r
    res, dt = timed(run, reps=2)
    report(f'C5 slant {ny}x{nx} on ONE GPU, 30 deg, GMAO-like cube {cfg["cube"]["wet"].shape}', ny * nx, dt, res,
           {'samples_per_ray': run.info.samples_per_ray, 'tiles': run.info.tiles, 'api': 'getInterpolators + _build_cube_ray'})

Path('gpurun_out').mkdir(exist_ok=True)
Path('gpurun_out/configs.json').write_text(json.dumps(out, indent=1))
