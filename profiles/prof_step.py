"""Short driver for ncu captures: two passes of the C2 hot path (K0 -> K3) and two K2 launches on materialised points.

    ncu --set full --clock-control none --import-source on -k regex:k_ray_integrate -c 1 -o gpurun_out/k3 python profiles/prof_step.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch  # noqa: E402

from bench import enu_const, global_config  # noqa: E402
from raider_b200 import _lib  # noqa: E402
from raider_b200.engine import DeviceCube  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
cfg = global_config(1)
xpts, ypts = cfg['xpts'][:n], cfg['ypts'][:n]
cube = DeviceCube.from_dict(cfg['cube'], device=0)
ow = torch.empty((n, n), dtype=torch.float64, device='cuda')
oh = torch.empty((n, n), dtype=torch.float64, device='cuda')
for _ in range(2):
    info = cube.trace(_lib.GEOM_GRID, xpts, ypts, n, n, _lib.LOS_ENU_CONST, enu_const(), 0.0, cfg['zref'], cfg['max_segment_length'], ow, oh)
nslots = int(sys.argv[2]) if len(sys.argv) > 2 else 16   # bench.py launches K2 on 48 slots (192 M points)
pts = torch.empty((nslots, n * n, 3), dtype=torch.float64, device='cuda')
cube.ray_points(info.maxlen, cfg['max_segment_length'], slot0=100, nslots=nslots, out=pts)
sw = torch.empty(nslots * n * n, dtype=torch.float64, device='cuda')
sh = torch.empty_like(sw)
for _ in range(2):
    cube.sample(pts.view(-1, 3), out=(sw, sh))
torch.cuda.synchronize()
print('samples/ray', info.samples_per_ray, 'checksum', float(ow.sum() + oh.sum()))
