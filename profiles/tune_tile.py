import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/profiles')
import numpy as np, torch
from bench import enu_const, global_config
from raider_b200 import _lib
from raider_b200.engine import DeviceCube
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')
def timed(fn, reps=7):
    out = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); torch.cuda.synchronize(); out.append(a.elapsed_time(b))
    return float(np.median(out))
cfg = global_config(1); enu = enu_const()
cube = DeviceCube.from_dict(cfg['cube'], device=0); cube.h.set_stream(stream.cuda_stream)
ny, nx = cfg['ypts'].size, cfg['xpts'].size
ow = torch.empty((ny, nx), dtype=torch.float64, device='cuda'); oh = torch.empty_like(ow)
maxlen, counts = cube.ray_layers(_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
ref = None
for tile in (0, 1, 2, 3, 4):
    for minb in (4,):
        os.environ['RDR_K3_TILE'] = str(tile); os.environ['RDR_K3_MINB'] = str(minb)
        t = timed(lambda: cube.ray_integrate(maxlen, cfg['max_segment_length'], False, ow, oh))
        w = ow.cpu().numpy()
        if ref is None: ref = w
        print(f'tile={tile} minb={minb}: K3 {t:.3f} ms  max|d| {np.abs(w-ref).max():.2e}', flush=True)
