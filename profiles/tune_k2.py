"""K2 variants: points per thread (RDR_K2_PPT) on materialised C2 sample points, f64 and f32 I/O."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from bench import enu_const, global_config  # noqa: E402
from raider_b200 import _lib  # noqa: E402
from raider_b200.engine import DeviceCube  # noqa: E402

cfg = global_config(1)
n = 2000
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
cube = DeviceCube.from_dict(cfg['cube'], device=0)
cube.h.set_stream(stream.cuda_stream)
maxlen, _ = cube.ray_layers(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], n, n, _lib.LOS_ENU_CONST, enu_const(), 0.0, cfg['zref'])
nslots = 48
pts = torch.empty((nslots, n * n, 3), dtype=torch.float64, device='cuda')
cube.ray_points(maxlen, cfg['max_segment_length'], slot0=100, nslots=nslots, out=pts)
npts = nslots * n * n
p32 = pts.to(torch.float32)
for dt, P, bpp, arith in ((torch.float64, pts, 40, 'f64'), (torch.float32, p32, 20, 'f64'), (torch.float32, p32, 20, 'f32')):
    sw = torch.empty(npts, dtype=dt, device='cuda')
    sh = torch.empty_like(sw)
    os.environ['RDR_K2_F32_ARITH'] = '1' if arith == 'f32' else '0'
    for ppt in ((1, 2, 4) if arith == 'f32' else (2,)):
        os.environ['RDR_K2_PPT'] = str(ppt)
        os.environ['RDR_K2_PPT32'] = str(ppt)
        for _ in range(3):
            cube.sample(P.view(-1, 3), out=(sw, sh))
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            cube.sample(P.view(-1, 3), out=(sw, sh))
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = float(np.median(ts))
        print(f'{str(dt):14s} arith={arith} ppt={ppt}: {t:.3f} ms  {npts * bpp / t / 1e6:.1f} GB/s  frac {npts * bpp / t / 1e6 / 6534.8:.3f}  checksum {float(sw.double().sum()):.6f}')
