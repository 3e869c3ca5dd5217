"""K3 variants on the C2 workload and on the 145-level table: time (CUDA events on the launching stream, L2 flushed) and
max |difference| against the PROJ-form integrator.

    python profiles/tune_k3.py
"""
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


cfg = global_config(1)
c145 = syn.config_c2(n=2000, table='ml145')
c145['xpts'], c145['ypts'] = cfg['xpts'], cfg['ypts']
variants = [('general', None, None, None), ('fast', 5, None, None), ('poly', None, None, None)] + [('poly', m, 1, q) for q in (0, 1) for m in (3, 4)] + [('poly', 4, 0, 0)]
for nm, cf in (('C2', cfg), ('ml145', c145)):
    cube = DeviceCube.from_dict(cf['cube'], device=0)
    cube.h.set_stream(stream.cuda_stream)
    ny, nx = cf['ypts'].size, cf['xpts'].size
    ow = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    oh = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    maxlen, counts = cube.ray_layers(_lib.GEOM_GRID, cf['xpts'], cf['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cf['zref'])
    ref = None
    for mode, minb, cache, quad in variants:
        os.environ['RDR_K3_MODE'] = mode
        for k, v in (('RDR_K3_MINB', minb), ('RDR_K3_CACHE', cache), ('RDR_K3_QUAD', quad)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        t3 = timed(lambda: cube.ray_integrate(maxlen, cf['max_segment_length'], False, ow, oh))
        w, h = ow.cpu().numpy(), oh.cpu().numpy()
        if ref is None:
            ref = (w, h)
        print(f'{nm} {mode} minb={minb} cache={cache} quad={quad}: K3 {t3:.3f} ms  max|d wet| {np.abs(w - ref[0]).max():.2e} max|d hydro| {np.abs(h - ref[1]).max():.2e} '
              f'fix {cube.h.last_fix_count}', flush=True)

# K0: Newton iterates on the span cubics of h(t) (default) vs on Bowring heights
for nm, cf, inc in (('C2', cfg, 30.0), ('ml145', c145, 30.0), ('C2 60deg', cfg, 60.0)):
    from raider_b200.losreader import inc_hd_to_enu
    e = np.ascontiguousarray(inc_hd_to_enu(np.float64(inc), np.float64(-168.0)))
    cube = DeviceCube.from_dict(cf['cube'], device=0)
    cube.h.set_stream(stream.cuda_stream)
    ny, nx = cf['ypts'].size, cf['xpts'].size
    ow = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    oh = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    ref = None
    for mode in ('exact', 'cubic'):
        os.environ['RDR_K0_MODE'] = mode
        for minb in (6, 8):
            os.environ['RDR_K0_MINB'] = str(minb)
            lay = lambda: cube.ray_layers(_lib.GEOM_GRID, cf['xpts'], cf['ypts'], ny, nx, _lib.LOS_ENU_CONST, e, 0.0, cf['zref'])
            maxlen, counts = lay()
            t0 = timed(lay)
            nparts, _ = cube.ray_integrate(maxlen, cf['max_segment_length'], False, ow, oh)
            w, h = ow.cpu().numpy(), oh.cpu().numpy()
            if ref is None:
                ref = (maxlen, w, h, nparts)
            print(f'{nm} K0 {mode} minb={minb}: {t0:.3f} ms  max|d maxlen| {np.abs(maxlen - ref[0]).max():.2e} m  nParts equal {np.array_equal(nparts, ref[3])}  '
                  f'max|d wet| {np.nanmax(np.abs(w - ref[1])):.2e} max|d hydro| {np.nanmax(np.abs(h - ref[2])):.2e}  nan {int(np.isnan(w).sum())}', flush=True)
    os.environ.pop('RDR_K0_MODE'); os.environ.pop('RDR_K0_MINB')
