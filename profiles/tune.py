"""Time K0 / K3 / K2 variants on the C2 workload (CUDA events on the launching stream, L2 flushed between launches).

    python profiles/tune.py            # all RDR_K3_MINB / RDR_K0_MINB variants
"""
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
ow = torch.empty((n, n), dtype=torch.float64, device='cuda')
oh = torch.empty((n, n), dtype=torch.float64, device='cuda')
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
enu = enu_const()


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


ref = None
layers = lambda: cube.ray_layers(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], n, n, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
for m in (4, 5, 6, 8):
    os.environ['RDR_K0_MINB'] = str(m)
    maxlen, _ = layers()
    print(f'K0 minBlocks={m}: {timed(layers):.3f} ms')
integ = lambda: cube.ray_integrate(maxlen, cfg['max_segment_length'], False, ow, oh)
for general in (0, 1):
    os.environ['RDR_K3_GENERAL'] = str(general)
    for npt in ((1, 2) if not general else (2,)):
        os.environ['RDR_K3_NPT'] = str(npt)
        for m in (3, 4, 5, 6, 8):
            os.environ['RDR_K3_MINB'] = str(m)
            integ()
            t3 = timed(integ)
            chk = float(ow.sum() + oh.sum())
            ref = ref or chk
            print(f'K3 {"general" if general else "fast"} npt={npt} minBlocks={m}: {t3:.3f} ms  checksum diff {chk - ref:.3e}')
os.environ['RDR_K3_GENERAL'] = '0'
for k in ('RDR_K3_NPT', 'RDR_K3_MINB', 'RDR_K0_MINB'):
    os.environ.pop(k, None)

nslots = 48
pts = torch.empty((nslots, n * n, 3), dtype=torch.float64, device='cuda')
cube.ray_points(maxlen, cfg['max_segment_length'], slot0=100, nslots=nslots, out=pts)
sw = torch.empty(nslots * n * n, dtype=torch.float64, device='cuda')
sh = torch.empty_like(sw)
t2 = timed(lambda: cube.sample(pts.view(-1, 3), out=(sw, sh)), reps=10)
gb = (nslots * n * n * 40) / 1e9
print(f'K2 stream f64: {t2:.3f} ms  {gb / t2 * 1e3:.1f} GB/s  ({gb / t2 * 1e3 / 6534.8:.3f} of measured peak)')
w_ref, h_ref = cube.sample(pts.view(-1, 3)[:200000].cpu().numpy())
print('K2 host-path == device-path:', np.array_equal(w_ref, sw[:200000].cpu().numpy(), equal_nan=True))
p32 = pts.to(torch.float32)
sw32 = torch.empty(nslots * n * n, dtype=torch.float32, device='cuda')
sh32 = torch.empty_like(sw32)
t2 = timed(lambda: cube.sample(p32.view(-1, 3), out=(sw32, sh32)), reps=10)
print(f'K2 stream f32 I/O: {t2:.3f} ms  {gb / 2 / t2 * 1e3:.1f} GB/s  ({gb / 2 / t2 * 1e3 / 6534.8:.3f} of measured peak)')
