"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-sharded fused step against the unsharded one.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/dist_check.py

Every rank traces its row block with the device-side exchange (K0 -> k_publish -> barrier -> k_plan -> K3 with peer stores ->
barrier); rank 0 also traces the WHOLE raster alone (no exchange) and compares bit for bit: maps, step counts, maxima.  Cases:
  a  C2-shaped raster, two output heights, the last one at the model top (that slice is skipped by the reference: zeros);
  b  the 145-node table (thin-layer kernel + quadrature kernel across ranks);
  c  a raster whose first rank's block has no look vectors at all (NaN): no rank may raise or hang, those rows come out NaN;
  d  the public API: build_cube_ray_sharded with host row blocks + device maps;
  e  66 deg incidence through the reference's 145-node table (80 km) at the default zref: every ray's last sample lies above max(z),
     the upper `.all()` clamp of delay.py:310-311 is decided over ALL ranks (its count rides the exchange slots); no NaN, bitwise equal.
Prints one JSON line on rank 0.
"""
import json
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from raider_b200 import _lib, synthetic as syn  # noqa: E402
from raider_b200.delay import _build_cube_ray  # noqa: E402
from raider_b200.delayFcns import getInterpolators  # noqa: E402
from raider_b200.dist import Comm, build_cube_ray_sharded, shard_rows  # noqa: E402
from raider_b200.losreader import Raytracing  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
comm = Comm()
res = {'world': world}


def same(a, b):
    return bool(np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True))


def case(name, cfg, los, zpts, **kw):
    ifs = getInterpolators(cfg['cube'], device=local)
    full = build_cube_ray_sharded(cfg['xpts'], cfg['ypts'], zpts, los, 4326, 4326, list(ifs), comm, MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                                  MAX_TROPO_HEIGHT=cfg['zref'], **kw)
    info = ifs[0].cube.last_info
    out = {'nan': int(np.isnan(np.asarray(full[0])).sum())}
    if rank == 0:
        solo_ifs = getInterpolators(cfg['cube'], device=local)
        want = _build_cube_ray(cfg['xpts'], cfg['ypts'], zpts, los, 4326, 4326, list(solo_ifs), MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                               MAX_TROPO_HEIGHT=cfg['zref'])
        winfo = solo_ifs[0].cube.last_info
        out.update(maps_bitwise_equal=same(full[0], want[0]) and same(full[1], want[1]),
                   max_abs_diff=float(np.nanmax(np.abs(np.asarray(full[0]) - want[0]))) if np.isfinite(want[0]).any() else 0.0,
                   nparts_equal=all(same(a.nparts, b.nparts) for a, b in zip(info, winfo) if not a.skipped),
                   maxlen_equal=all(same(a.maxlen, b.maxlen) for a, b in zip(info, winfo) if not a.skipped),
                   skipped=[bool(a.skipped) for a in info], k_split=[int(a.k_split) for a in info],
                   clamp_high_last=[bool(a.clamp_high_last) for a in info])
    res[name] = out


# a: two heights, the last at the model top
cfg = syn.config_c2(n=96)
top = float(cfg['cube']['z'][-1])
case('a_two_heights_top_slice_skipped', cfg, Raytracing(incidence=30.0, heading=-168.0), np.array([0.0, top]))
if rank == 0:
    res['a_two_heights_top_slice_skipped']['top_slice_all_zero'] = None  # filled below from the gathered maps
ifs = getInterpolators(cfg['cube'], device=local)
full = build_cube_ray_sharded(cfg['xpts'], cfg['ypts'], np.array([0.0, top]), Raytracing(incidence=30.0, heading=-168.0), 4326, 4326, list(ifs), comm,
                              MAX_SEGMENT_LENGTH=225.0, MAX_TROPO_HEIGHT=cfg['zref'])
zero_ok = torch.tensor([int((np.asarray(full[0])[1] == 0).all() and (np.asarray(full[1])[1] == 0).all())], device='cuda')
dist.all_reduce(zero_ok, op=dist.ReduceOp.MIN)
res['a_two_heights_top_slice_skipped']['top_slice_all_zero_on_every_rank'] = bool(zero_ok.item())

# b: 145-node table
cfg = syn.config_c2(n=128, table='ml145')
case('b_ml145', cfg, Raytracing(incidence=37.0, heading=15.0), np.array([0.0]))

# c: the first rank's rows have NaN look vectors
cfg = syn.config_c2(n=64)
xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
from raider_b200.losreader import inc_hd_to_enu  # noqa: E402
from raider_b200.utilFcns import enu2ecef  # noqa: E402
enu = inc_hd_to_enu(np.float64(30.0), np.float64(-168.0))
vec = enu2ecef(enu[0], enu[1], enu[2], yy, xx, 0 * yy)
r0, r1 = shard_rows(64, 0, world)
vec[r0:r1] = np.nan

class BlockLOS(Raytracing):
    """look vectors handed out for whatever row block is asked for"""
    def __init__(self, vecs, ypts):
        super().__init__(look_vecs=vecs)
        self._all, self._ypts = vecs, ypts

    def getLookVectors(self, ht, llh, xyz, yy):
        rows = np.searchsorted(-self._ypts, -np.asarray(yy)[:, 0])
        return np.ascontiguousarray(self._all[rows])


case('c_first_block_all_nan', cfg, BlockLOS(vec, cfg['ypts']), np.array([0.0]))
res['c_first_block_all_nan']['expected_nan'] = int((r1 - r0) * 64)

# d: host row blocks + device maps
cfg = syn.config_c2(n=96)
ifs = getInterpolators(cfg['cube'], device=local)
dev_maps, host_rows = build_cube_ray_sharded(cfg['xpts'], cfg['ypts'], np.array([0.0]), Raytracing(incidence=30.0, heading=-168.0), 4326, 4326, list(ifs), comm,
                                             MAX_SEGMENT_LENGTH=225.0, MAX_TROPO_HEIGHT=cfg['zref'], gather='device', host_block=True)
r0, r1 = shard_rows(96, rank, world)
ok = torch.tensor([int(same(host_rows[0][0], dev_maps[0][0, r0:r1].cpu().numpy()) and same(host_rows[1][0], dev_maps[1][0, r0:r1].cpu().numpy()))], device='cuda')
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
res['d_host_rows_equal_device_maps_on_every_rank'] = bool(ok.item())

# e: the upper clamp of the last sample, decided globally
zs = np.load(Path(__file__).resolve().parents[1] / 'tests' / 'golden' / 'era5_slant_ref.npz')['z']
xp, yp = syn.raster(33.5, -117.8, 64, 64, 0.02)
xs, ys = syn.cube_axes_around(xp, yp, pad_deg=3.0)
cfg = {'cube': syn.make_cube(ys, xs, zs, totals=False), 'xpts': xp, 'ypts': yp, 'zref': float(zs[-1] - 1), 'max_segment_length': 1000.0}
case('e_upper_clamp_66deg', cfg, Raytracing(incidence=66.0, heading=-168.0), np.array([0.0]))

if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
