"""Round-2 kernel check: the fused step (K0 -> device plan -> K3 quadrature + thin-layer kernels) on C2 and on the 145-node table
at the reference's 1000 m segments; per-stage CUDA-event timings (L2 flushed) and max |difference| against the PROJ-form
integrator on the same rays.

    python profiles/r02_check.py [c2|ml145|hrrr57 ...]
"""
import json
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


def configs(names):
    base = global_config(1)
    for nm in names:
        if nm == 'c2':
            yield nm, base
        else:
            c = syn.config_c2(n=2000, table=nm)
            c['xpts'], c['ypts'] = base['xpts'], base['ypts']
            yield nm, c


res = {}
for nm, cf in configs(sys.argv[1:] or ['c2', 'ml145']):
    cube = DeviceCube.from_dict(cf['cube'], device=0)
    cube.h.set_stream(stream.cuda_stream)
    ny, nx = cf['ypts'].size, cf['xpts'].size
    ow = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    oh = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    args = (_lib.GEOM_GRID, cf['xpts'], cf['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cf['zref'])
    S = cf['max_segment_length']
    for _ in range(2):
        info = cube.trace(*args, S, ow, oh)
    t_step = timed(lambda: (cube.trace_begin(*args), cube.trace_finish(S, ow, oh)))
    t_k0 = timed(lambda: cube.trace_begin(*args))
    cube.trace_begin(*args)
    t_k3 = timed(lambda: cube.trace_finish(S, ow, oh))
    w, h = ow.cpu().numpy(), oh.cpu().numpy()
    # PROJ-form integrator on the same rays as the yardstick
    cube.trace_begin(*args)
    cube.trace_finish(S, ow, oh, mode=_lib.K3_GENERAL)
    torch.cuda.synchronize()
    wg, hg = ow.cpu().numpy(), oh.cpu().numpy()
    r = {'ms_step': t_step, 'ms_k0': t_k0, 'ms_k3': t_k3, 'rays_per_s': ny * nx / (t_step * 1e-3), 'layers': info.n_layers,
         'samples_per_ray': info.samples_per_ray, 'k_split': info.k_split, 'n_spans': info.n_spans, 'staged': [info.staged_passes, info.unstaged_passes],
         'nparts_hist': np.bincount(info.nparts).tolist(), 'max_abs_diff_vs_general_m': float(max(np.abs(w - wg).max(), np.abs(h - hg).max())),
         'nan': int(np.isnan(w).sum()), 'fix_count': cube.h.last_fix_count, 'env': {k: v for k, v in os.environ.items() if k.startswith('RDR_')}}
    res[nm] = r
    print(nm, json.dumps(r), flush=True)
