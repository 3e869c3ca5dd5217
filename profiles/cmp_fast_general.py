"""Polynomial / fast vs PROJ-form integrator on the C2 workload (and an oblique / polar / partially-outside variant): max |difference|,
rays handed to the fix-up pass, kernel times (CUDA events on the launching stream, L2 flushed)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from bench import global_config  # noqa: E402
from raider_b200 import _lib, synthetic as syn  # noqa: E402
from raider_b200.engine import DeviceCube  # noqa: E402
from raider_b200.losreader import inc_hd_to_enu  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')


def timed(fn, reps=5):
    out = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return float(np.median(out))


def run(name, cfg, inc, n, time_it=True):
    cube = DeviceCube.from_dict(cfg['cube'], device=0)
    cube.h.set_stream(stream.cuda_stream)
    enu = np.ascontiguousarray(inc_hd_to_enu(np.float64(inc), np.float64(-168.0)))
    ny, nx = cfg['ypts'].size, cfg['xpts'].size
    res = {}
    for mode in ('poly', 'fast', 'general'):
        os.environ['RDR_K3_MODE'] = mode
        ow = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
        oh = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
        t0 = timed(lambda: cube.ray_layers(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref']))
        maxlen, counts = cube.ray_layers(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
        clamp = bool(counts[2] == counts[0])
        t3 = timed(lambda: cube.ray_integrate(maxlen, cfg['max_segment_length'], clamp, ow, oh)) if time_it else float('nan')
        nparts, oob = cube.ray_integrate(maxlen, cfg['max_segment_length'], clamp, ow, oh)
        res[mode] = (ow.cpu().numpy(), oh.cpu().numpy(), t0, t3, cube.h.last_fix_count, int(nparts.sum()), oob)
    g = res['general']
    line = f'{name}: rays {ny * nx}, samples/ray {g[5]}, NaN rays {int(np.isnan(g[0]).sum())}, K0 {g[2]:.3f} ms, K3 general {g[3]:.3f} ms'
    for mode in ('poly', 'fast'):
        f = res[mode]
        both_nan = np.isnan(f[0]) == np.isnan(g[0])
        dw = np.nanmax(np.abs(f[0] - g[0])) if np.isfinite(g[0]).any() else 0.0
        dh = np.nanmax(np.abs(f[1] - g[1])) if np.isfinite(g[1]).any() else 0.0
        line += (f'\n    {mode}: K3 {f[3]:.3f} ms, fixed-up rays {f[4]}, NaN pattern equal {bool(both_nan.all())}, max|{mode}-general| wet {dw:.3e} hydro {dh:.3e} m, '
                 f'oob {f[6]} vs {g[6]}')
    print(line, flush=True)
    os.environ.pop('RDR_K3_MODE', None)


n = int(os.environ.get('N', '2000'))
cfg = global_config(1)
if n != 2000:
    cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, n, n, 0.001 * 2000 / n)
run('C2 30deg', cfg, 30.0, n)
run('C2 45deg', cfg, 45.0, n)
run('C2 70deg (oblique: Newton tops metres off, some rays leave the window)', cfg, 70.0, n)
# table variant: 145 levels, 1000 m segments
c145 = syn.config_c2(n=n, table='ml145')
c145['xpts'], c145['ypts'] = cfg['xpts'], cfg['ypts']
run('C2 ml145 30deg', c145, 30.0, n)
# polar: raster centred at 85N
m = min(n, 400)
xp, yp = syn.raster(85.0, 20.0, m, m, 0.002)
xs, ys = syn.cube_axes_around(xp, yp)
pol = {'cube': syn.make_cube(ys, xs, syn.z_levels(37), totals=False), 'xpts': xp, 'ypts': yp, 'zref': cfg['zref'], 'max_segment_length': 225.0}
run('polar 85N', pol, 30.0, m, time_it=False)
# raster hanging over the cube edge: pad only 0.05 deg so slanted rays leave the cube
xp, yp = syn.raster(34.0, -118.0, m, m, 0.004)
xs, ys = syn.cube_axes_around(xp, yp, pad_deg=0.0)
edge = {'cube': syn.make_cube(ys, xs, syn.z_levels(37), totals=False), 'xpts': xp, 'ypts': yp, 'zref': cfg['zref'], 'max_segment_length': 225.0}
run('edge (rays leave the cube)', edge, 30.0, m, time_it=False)

# span / occupancy sweep of the polynomial integrator on C2 and on the 145-level table
for nm, cf in (('C2', cfg), ('ml145', c145)):
    cube = DeviceCube.from_dict(cf['cube'], device=0)
    cube.h.set_stream(stream.cuda_stream)
    enu = np.ascontiguousarray(inc_hd_to_enu(np.float64(30.0), np.float64(-168.0)))
    ny, nx = cf['ypts'].size, cf['xpts'].size
    ow = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    oh = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    maxlen, counts = cube.ray_layers(_lib.GEOM_GRID, cf['xpts'], cf['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cf['zref'])
    os.environ['RDR_K3_MODE'] = 'general'
    cube.ray_integrate(maxlen, cf['max_segment_length'], False, ow, oh)
    ref = ow.cpu().numpy() + oh.cpu().numpy()
    os.environ['RDR_K3_MODE'] = 'poly'
    for span in (2000, 4000, 8000, 12000, 16000, 24000):
        for minb in (3, 4, 5):
            os.environ['RDR_K3_SPAN'] = str(span)
            os.environ['RDR_K3_MINB'] = str(minb)
            t3 = timed(lambda: cube.ray_integrate(maxlen, cf['max_segment_length'], False, ow, oh))
            d = float(np.abs(ow.cpu().numpy() + oh.cpu().numpy() - ref).max())
            print(f'{nm} poly span {span} minb {minb}: K3 {t3:.3f} ms, max|total - general| {d:.3e} m', flush=True)
    os.environ.pop('RDR_K3_SPAN'); os.environ.pop('RDR_K3_MINB'); os.environ.pop('RDR_K3_MODE')
