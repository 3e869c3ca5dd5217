"""K3 (polynomial + layer quadrature) vs span length and occupancy on C2; max |difference| against the PROJ-form integrator."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from bench import enu_const, global_config  # noqa: E402
from raider_b200 import _lib  # noqa: E402
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
cube = DeviceCube.from_dict(cfg['cube'], device=0)
cube.h.set_stream(stream.cuda_stream)
ny, nx = cfg['ypts'].size, cfg['xpts'].size
ow = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
oh = torch.empty_like(ow)
maxlen, counts = cube.ray_layers(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
os.environ['RDR_K3_MODE'] = 'general'
cube.ray_integrate(maxlen, cfg['max_segment_length'], False, ow, oh)
ref = (ow.cpu().numpy(), oh.cpu().numpy())
os.environ['RDR_K3_MODE'] = 'poly'
for span in (6000, 8000, 12000, 16000, 24000):
    for minb in (3, 4):
        os.environ['RDR_K3_SPAN'] = str(span)
        os.environ['RDR_K3_MINB'] = str(minb)
        t = timed(lambda: cube.ray_integrate(maxlen, cfg['max_segment_length'], False, ow, oh))
        d = max(np.abs(ow.cpu().numpy() - ref[0]).max(), np.abs(oh.cpu().numpy() - ref[1]).max())
        print(f'span {span} minb {minb}: K3 {t:.3f} ms  max|d| {d:.2e} m', flush=True)
