"""K3 time on the C2 rays as a function of the horizontal cell size of the cube: how much of the kernel goes into layers that
cross a horizontal cell face (summed sample by sample) instead of the closed-form layer sum."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from bench import enu_const, global_config  # noqa: E402
from raider_b200 import _lib, synthetic as syn  # noqa: E402
from raider_b200.engine import DeviceCube  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')


def timed(fn, reps=7):
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


cfg = global_config(1)
enu = enu_const()
ny, nx = cfg['ypts'].size, cfg['xpts'].size
ow = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
oh = torch.empty_like(ow)
for spacing in (0.0625, 0.125, 0.25, 0.5, 1.0, 4.0):
    xs, ys = syn.cube_axes_around(cfg['xpts'], cfg['ypts'], spacing_deg=spacing, pad_deg=max(2.0, 2 * spacing))
    cube = DeviceCube.from_dict(syn.make_cube(ys, xs, syn.z_levels(37), totals=False), device=0)
    cube.h.set_stream(stream.cuda_stream)
    maxlen, counts = cube.ray_layers(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
    res = {}
    for quad in (1, 0):
        os.environ['RDR_K3_QUAD'] = str(quad)
        res[quad] = timed(lambda: cube.ray_integrate(maxlen, cfg['max_segment_length'], False, ow, oh))
    print(f'cells {spacing} deg ({ys.size} x {xs.size}): K3 quadrature {res[1]:.3f} ms, per-sample {res[0]:.3f} ms', flush=True)
