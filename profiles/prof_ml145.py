"""ncu driver: the C2 rays through the 145-level table at the reference's default 1000 m segments (the production setting)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from bench import enu_const, global_config  # noqa: E402
from raider_b200 import _lib, synthetic as syn  # noqa: E402
from raider_b200.engine import DeviceCube  # noqa: E402

cfg = global_config(1)
c145 = syn.config_c2(n=2000, table='ml145')
c145['xpts'], c145['ypts'] = cfg['xpts'], cfg['ypts']
n = 2000
cube = DeviceCube.from_dict(c145['cube'], device=0)
ow = torch.empty((n, n), dtype=torch.float64, device='cuda')
oh = torch.empty((n, n), dtype=torch.float64, device='cuda')
for _ in range(2):
    info = cube.trace(_lib.GEOM_GRID, c145['xpts'], c145['ypts'], n, n, _lib.LOS_ENU_CONST, enu_const(), 0.0, c145['zref'], c145['max_segment_length'], ow, oh)
torch.cuda.synchronize()
print('samples/ray', info.samples_per_ray, 'layers', info.n_layers, 'nparts', np.bincount(info.nparts), 'checksum', float(ow.sum() + oh.sum()))
