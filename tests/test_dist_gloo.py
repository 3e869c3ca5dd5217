"""N > 1 path on CPU: world_size-2 `gloo` process group drives raider_b200.dist (row sharding, the all-reduce(MAX) of the
per-layer maxima, the all-reduce(SUM) of the predicate counters, the all-gather of the delay maps) with the oracle standing
in for the kernels.  Sharded == unsharded only because the maxima are reduced globally."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _oracle_build_fn(xpts, ypts, zpts, los, model_crs, pts_crs, interpolators, MAX_SEGMENT_LENGTH, MAX_TROPO_HEIGHT, reduce_max, reduce_sum):
    """Stand-in for the device path: local K0 (build_ray maxima) -> global reduction -> local integration with the global nParts."""
    from oracle import geodesy, raytrace as rt
    model_zs = interpolators[0].grid[2]
    xx, yy = np.meshgrid(xpts, ypts)
    maxima = []
    for ht in zpts:
        xyz = np.stack(geodesy.lla2ecef(yy, xx, np.full(yy.shape, ht)), -1)
        look = los.getLookVectors(ht, [xx, yy, np.full(yy.shape, ht)], xyz, yy)
        lens = rt.build_ray(model_zs, ht, xyz, look, MAX_TROPO_HEIGHT)[0]
        maxima.append(reduce_max(lens.max((1, 2))))
        assert int(reduce_sum(np.array([yy.size]))[0]) > yy.size  # the counters really are summed over ranks
    return rt.build_cube_ray(xpts, ypts, zpts, los, model_crs, pts_crs, interpolators, MAX_SEGMENT_LENGTH=MAX_SEGMENT_LENGTH,
                             MAX_TROPO_HEIGHT=MAX_TROPO_HEIGHT, layer_maxlen=maxima)


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import geodesy, raytrace as rt
    from raider_b200 import synthetic as syn
    from raider_b200.dist import Comm, build_cube_ray_sharded, shard_rows
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        comm = Comm()
        cfg = syn.config_c2(n=13)  # odd row count: uneven blocks (7 + 6)
        cfg['xpts'], cfg['ypts'] = syn.raster(34.0, -118.0, 13, 9, 0.14)
        xx, yy = np.meshgrid(cfg['xpts'], cfg['ypts'])
        inc = 18.0 + 30.0 * (yy - yy.min()) / (yy.max() - yy.min())  # incidence grows with the row: per-layer maxima differ per shard
        enu = geodesy.inc_hd_to_enu(inc, np.full(inc.shape, -168.0))
        vecs = geodesy.enu2ecef(enu[..., 0], enu[..., 1], enu[..., 2], yy, xx, 0 * yy)

        class RowLOS:  # look vectors by row, whichever block of rows is asked for
            def getLookVectors(self, ht, llh, xyz, yy_blk):
                rows = [int(np.argmin(np.abs(cfg['ypts'] - y))) for y in yy_blk[:, 0]]
                return vecs[rows]

        crs = rt.GeographicCRS()
        ifs = list(rt.get_interpolators(cfg['cube']))
        zpts = np.array([0.0, 800.0])
        out = build_cube_ray_sharded(cfg['xpts'], cfg['ypts'], zpts, RowLOS(), crs, crs, ifs, comm, MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                                     MAX_TROPO_HEIGHT=cfg['zref'], build_fn=_oracle_build_fn)
        full = rt.build_cube_ray(cfg['xpts'], cfg['ypts'], zpts, rt.ArrayLOS(vecs), crs, crs, ifs, MAX_SEGMENT_LENGTH=cfg['max_segment_length'],
                                 MAX_TROPO_HEIGHT=cfg['zref'])
        r0, r1 = shard_rows(13, rank, world)
        naive = rt.build_cube_ray(cfg['xpts'], cfg['ypts'][r0:r1], zpts, rt.ArrayLOS(vecs[r0:r1]), crs, crs, ifs,
                                  MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
        gathered = comm.all_gather_rows(np.full((r1 - r0, 4), float(rank)), 13)
        # both reductions of a step in one collective (what engine._global_plan uses when the hooks come from a Comm)
        from raider_b200.engine import _global_plan
        pm, pc = comm.reduce_pair(np.array([float(rank), 5.0 - rank, 2.5]), np.array([rank + 1, 10, 0]))
        # counts of rdr_ray_layers: {rays, NaN rays, first samples below min(z), K, last samples above max(z)}: all but K are summed
        gm, gc = _global_plan(np.array([float(rank), 5.0 - rank, 2.5]), np.array([rank + 1, 10, 0, 33, rank + 1], dtype=np.int64), comm.reduce_max, comm.reduce_sum)
        pair_ok = pm.tolist() == [1.0, 5.0, 2.5] and pc.tolist() == [3, 20, 0] and gm.tolist() == pm.tolist() and gc.tolist() == [3, 20, 0, 33, 3]
        q.put((rank, bool(np.array_equal(out[0], full[0]) and np.array_equal(out[1], full[1])), out[0].shape,
               bool(np.array_equal(naive[0], full[0][:, r0:r1])), gathered[:, 0].tolist(),
               comm.reduce_max(np.array([float(rank), 5.0 - rank])).tolist(), comm.reduce_sum(np.array([rank + 1, 10])).tolist(), pair_ok))
    finally:
        dist.destroy_process_group()


def test_shard_rows_partition():
    from raider_b200.dist import shard_rows
    for ny in (1, 2, 7, 13, 2000, 16001):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_rows(ny, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == ny
            assert all(b[1] == blocks[i + 1][0] for i, b in enumerate(blocks[:-1]))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_world_size_2_gloo_sharded_equals_unsharded():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, shape, naive_same, gathered, rmax, rsum, pair_ok in res:
        assert pair_ok, 'reduce_pair / _global_plan disagree with the separate MAX and SUM reductions'
        assert same, 'sharded result differs from the unsharded oracle although the maxima were reduced globally'
        assert shape == (2, 13, 9)
        assert gathered == [0.0] * 7 + [1.0] * 6          # uneven row blocks reassembled in order
        assert rmax == [1.0, 5.0] and rsum == [3, 20]
    # without the global reduction at least one shard integrates with different step counts -> different numbers
    assert not all(r[3] for r in res)
