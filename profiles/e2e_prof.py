import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from bench import global_config, INC, HEAD
from raider_b200.delay import _build_cube_ray
from raider_b200.delayFcns import getInterpolators
from raider_b200.losreader import Raytracing
cfg = global_config(1)
los = Raytracing(incidence=INC, heading=HEAD)
def step():
    t0 = time.perf_counter()
    ifs = getInterpolators(cfg['cube'], device=0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    r = _build_cube_ray(cfg['xpts'], cfg['ypts'], np.array([0.0]), los, 4326, 4326, list(ifs), MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3
for _ in range(3): step()
a = np.array([step() for _ in range(10)])
print('getInterpolators %.3f ms, _build_cube_ray %.3f ms' % tuple(np.median(a, 0)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
