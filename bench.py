#!/usr/bin/env python
"""bench.py -- LOS rays/s of the slant-delay hot path on BASELINE.json config C2, with the roofline and CPU baseline beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic input: the C2 raster (2000 x 2000 rays per GPU, fixed
30 deg incidence, NZ = 37 cube, 225 m segments => ~296 samples/ray, one output height):  K0 ray_layers -> global
max/predicate reduction -> K3 ray_integrate; when N > 1, K3 also stores every ray into the full maps of all GPUs (peer-mapped
symmetric memory over NVLink: the all-gather of the output maps, fused into the kernel; NCCL all-gather as the fallback).

One JSON line on stdout (rank 0):
  value     rays/s, whole job, geometry + cube resident in HBM, outputs left in HBM, CUDA events, max over ranks
  e2e       rays/s through the reference-facing API (getInterpolators + _build_cube_ray) with HOST buffers: cube H2D,
            axes H2D, both delay maps D2H inside the timed region
  roofline  the unfused trilinear-sample kernel K2 on materialised sample points of the same rays (40 B/point fp64),
            achieved HBM GB/s vs MEASURED_PEAKS.json -- the kernel north_star puts the HBM-roofline claim on; with the fp32
            tier (20 B/point) and the CPU samplers (scipy RGI, the reference's native interpolate) on a bounded sample beside it
  fused     the K3 kernel's own byte accounts (it is fp64-issue bound, not HBM bound)
  cpu_baseline  the oracle port (NumPy + scipy restatement of the reference loops) on a bounded sub-raster, 1 core

--impl reference times that same CPU port on all host cores (row blocks in worker processes, global nParts injected);
the reference itself is single-process (delay.py:133,178-185) and cannot be imported offline (pyproj/xarray/isce3).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'LOS rays/sec, slant delay (ray-traced), C2: 2000x2000 raster, 30deg incidence, NZ=37 cube, ~300 steps/ray'
UNIT = 'rays/s'
N_SIDE = 2000
INC, HEAD = 30.0, -168.0


# ----------------------------------------------------------------------------------------------------------------
def measured_peak_gbs():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        try:
            return float(json.loads(p.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms during the timed regions (B200_PROFILING.md): the device-resident
    steps, the end-to-end steps and the K2 / K3 kernel timings all run inside it."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int) -> None:
        self.rows, self.proc, self.idx = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '50', '-i', str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, reasons, smax = [], set(), None
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(nme)
            except (ValueError, IndexError):
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(sm)}


def global_config(n_gpus: int):
    """Weak scaling: every GPU owns a 2000 x 2000 block of rows; the global raster is (2000 N) x 2000 at 0.001 deg."""
    from raider_b200 import synthetic as syn
    ny = N_SIDE * n_gpus
    xpts, ypts = syn.raster(34.0, -118.0, ny, N_SIDE, 0.001)
    xs, ys = syn.cube_axes_around(xpts, ypts)
    zs = syn.z_levels(37)
    cube = syn.make_cube(ys, xs, zs, totals=False)
    return {'cube': cube, 'xpts': xpts, 'ypts': ypts, 'zref': float(zs[-1] - 1.0), 'max_segment_length': 225.0}


def c5_config(table=None):
    """BASELINE configs[4]: NISAR-scale raster, 12000 x 20000 = 2.4e8 rays, GMAO-like cube 0.25 x 0.3125 deg, NZ = 72 (or the
    145-node table), fixed 30 deg incidence, the reference's default 1000 m segments; STRONG scaling: the rows are split over the GPUs."""
    from raider_b200 import synthetic as syn
    return syn.config_c5(table=table)


def workload_config(world: int, cube: dict, config: str = 'c2') -> dict:
    """The `config` both arms print."""
    if config == 'c5':
        return {'workload': f'C5 NISAR-scale slant delay: 12000x20000 = 2.4e8 rays over {world} GPU(s) (strong scaling), fixed {INC} deg incidence, '
                            f'heading {HEAD}, 0.0002 deg posting, cube {cube["y"].size}x{cube["x"].size}x{cube["z"].size} @0.25x0.3125 deg fp32, 1000 m max segment',
                'rays_per_step': 12000 * 20000}
    return {'workload': f'C2 slant delay: {N_SIDE}x{N_SIDE} rays per GPU ({N_SIDE * world}x{N_SIDE} global), fixed {INC} deg incidence, heading {HEAD}, '
                        f'0.001 deg posting, cube {cube["y"].size}x{cube["x"].size}x37 @0.25 deg fp32, 225 m max segment',
            'rays_per_step': N_SIDE * world * N_SIDE}


def kernel_source_hash() -> str:
    """sha256 over the kernel sources + the header: profiles/k2_traffic.json is only trusted when it was captured from these."""
    import hashlib
    h = hashlib.sha256()
    for p in sorted((ROOT / 'raider_b200' / 'csrc').glob('*')) + [ROOT / 'include' / 'raider_b200.h']:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()[:16]


def enu_const():
    from raider_b200.losreader import inc_hd_to_enu
    return np.ascontiguousarray(inc_hd_to_enu(np.float64(INC), np.float64(HEAD)))


# ----------------------------------------------------------------------------------------------------------------
# CPU arms (oracle port)
# ----------------------------------------------------------------------------------------------------------------
def _cpu_block(args):
    """One worker: the oracle's _build_cube_ray on a row block with the global per-layer maxima injected."""
    cube, xpts, ypts, zref, seg, maxlen = args[:6]
    want_out = len(args) > 6 and args[6]
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from oracle import raytrace as rt
    crs = rt.GeographicCRS()
    t0 = time.perf_counter()
    out = rt.build_cube_ray(xpts, ypts, np.array([0.0]), rt.FixedIncidenceLOS(INC, HEAD), crs, crs, list(rt.get_interpolators(cube)),
                            MAX_SEGMENT_LENGTH=seg, MAX_TROPO_HEIGHT=zref, layer_maxlen=None if maxlen is None else [maxlen])
    dt = time.perf_counter() - t0
    return (dt, out) if want_out else (dt, float(out[0].sum() + out[1].sum()))


def cpu_global_maxlen(cfg, interior_step: int = 20):
    """The ORACLE'S OWN per-layer maxima of the full raster (delay.py:283) -- nothing borrowed from the GPU: losreader.build_ray's
    restatement on the border pixels plus every `interior_step`-th row and column of the interior (the ray length is a smooth
    function of position, its maximum sits on the border; the interior sample is the check of that).  Returns (maxima, border
    maxima): the two must coincide."""
    from oracle import geodesy, raytrace as rt
    xp, yp = cfg['xpts'], cfg['ypts']
    bx = np.concatenate([xp, xp, np.full(yp.size, xp[0]), np.full(yp.size, xp[-1])])
    by = np.concatenate([np.full(xp.size, yp[0]), np.full(xp.size, yp[-1]), yp, yp])
    ix, iy = np.meshgrid(xp[::interior_step], yp[::interior_step])
    los = rt.FixedIncidenceLOS(INC, HEAD)

    def maxima(xx, yy):
        xx, yy = xx.reshape(1, -1), yy.reshape(1, -1)
        xyz = np.stack(geodesy.lla2ecef(yy, xx, np.zeros_like(yy)), -1)
        look = los.getLookVectors(0.0, [xx, yy, 0 * yy], xyz, yy)
        return rt.build_ray(cfg['cube']['z'], 0.0, xyz, look, cfg['zref'])[0].max((1, 2))
    border, interior = maxima(bx, by), maxima(ix, iy)
    return np.maximum(border, interior), border


def cpu_baseline_single(cfg, maxlen, rows=100, want_out=False):
    """cpu_baseline leg: 1 core, `rows` x 2000 rays from the middle of the raster."""
    mid = cfg['ypts'].size // 2
    dt, out = _cpu_block((cfg['cube'], cfg['xpts'], cfg['ypts'][mid:mid + rows], cfg['zref'], cfg['max_segment_length'], maxlen, want_out))
    n = rows * cfg['xpts'].size
    res = {'value': n / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
           'sample': f'{rows}x{cfg["xpts"].size} rays (rows {mid}..{mid + rows - 1} of the raster), nParts from the oracle\'s own full-raster maxima, {dt:.1f} s'}
    return (res, out, slice(mid, mid + rows)) if want_out else res


def run_reference(args):
    """--impl reference: the CPU port on all host cores; each step = cores x rows_per_proc x 2000 rays of the C2 raster."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cfg = global_config(1)
    cores = os.cpu_count() or 1
    maxlen, _ = cpu_global_maxlen(cfg)
    # calibrate: ~6 s per step so that (steps + warmup) stays within a few minutes
    t_probe, _ = _cpu_block((cfg['cube'], cfg['xpts'], cfg['ypts'][1000:1002], cfg['zref'], cfg['max_segment_length'], maxlen))
    budget = min(8.0, 150.0 / max(1, args.steps + args.warmup))
    rows = int(max(1, min(N_SIDE // cores, round(budget / (t_probe / 2.0)))))
    blocks = [(cfg['cube'], cfg['xpts'], cfg['ypts'][i * rows:(i + 1) * rows], cfg['zref'], cfg['max_segment_length'], maxlen) for i in range(cores)]
    n_step = cores * rows * N_SIDE
    with mp.get_context('fork').Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_cpu_block, blocks)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_block, blocks)
        dt = time.perf_counter() - t0
    value = n_step * args.steps / dt
    sample = f'{cores} processes x {rows} rows x {N_SIDE} rays per step of the C2 raster, global nParts injected'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': dict(workload_config(1, cfg['cube']), reference_sample_rays_per_step=n_step,
                                            note='each step integrates a bounded sample of the workload (rows of the same raster, same cube, '
                                                 'same samples per ray); rays/s is per ray actually integrated'),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def _events(torch, stream, flush, fn, reps):
    out = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return float(np.mean(out))


def run_variant(torch, stream, flush, name, cfg, enu, model_crs=None, reps=5):
    """One more workload beside the headline, on one GPU: the fused step (K0 -> device plan -> K3) with CUDA events around the
    whole step and around its two halves, and the difference to the PROJ-form integrator on the same rays."""
    from raider_b200 import _lib
    from raider_b200.engine import DeviceCube
    cube = DeviceCube.from_dict(cfg['cube'], crs=model_crs)
    cube.h.set_stream(stream.cuda_stream)
    ny, nx = cfg['ypts'].size, cfg['xpts'].size
    ow = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    oh = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    a = (_lib.GEOM_GRID, cfg['xpts'], cfg['ypts'], ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
    S = cfg['max_segment_length']
    for _ in range(3):
        info = cube.trace(*a, S, ow, oh)
    ms_step = _events(torch, stream, flush, lambda: cube.trace(*a, S, ow, oh), reps)
    ms_k0 = _events(torch, stream, flush, lambda: cube.trace_begin(*a), reps)
    cube.trace_begin(*a)
    ms_k3 = _events(torch, stream, flush, lambda: cube.trace_finish(S, ow, oh), reps)
    w = ow[::40].clone()
    cube.trace_begin(*a)
    cube.trace_finish(S, ow, oh, mode=_lib.K3_GENERAL)
    torch.cuda.synchronize()
    diff = float((ow[::40] - w).abs().max().item())
    del cube, ow, oh
    return {'workload': name, 'rays': ny * nx, 'rays_per_s': ny * nx / (ms_step * 1e-3), 'ms_per_step': ms_step, 'ms_k0': ms_k0, 'ms_k3': ms_k3,
            'layers': info.n_layers, 'samples_per_ray': info.samples_per_ray, 'layers_in_thin_kernel': info.k_split,
            'tma_staged_passes': [info.staged_passes, info.unstaged_passes], 'max_abs_diff_vs_proj_form_m': diff, 'nan': int(torch.isnan(w).sum().item())}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from raider_b200 import _lib, synthetic as syn
    from raider_b200.delay import _build_cube_ray
    from raider_b200.delayFcns import getInterpolators
    from raider_b200.dist import Comm, shard_rows
    from raider_b200.engine import DeviceCube
    from raider_b200.losreader import Raytracing

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f'--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}')
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        comm = Comm()

    c5 = args.config.startswith('c5')
    if c5:
        cfg = c5_config('ml145' if args.config == 'c5-ml145' else None)
        metric = METRIC.replace('C2: 2000x2000 raster, 30deg incidence, NZ=37 cube, ~300 steps/ray',
                                'C5: 12000x20000 raster, 30deg incidence, ' + ('145-node table' if args.config == 'c5-ml145' else 'NZ=72 cube') + ', 1000 m segments')
    else:
        cfg = global_config(world)
        metric = METRIC
    ny_g, nx = cfg['ypts'].size, cfg['xpts'].size
    r0, r1 = shard_rows(ny_g, rank, world)
    ypts = np.ascontiguousarray(cfg['ypts'][r0:r1])
    ny = ypts.size
    n_local, n_global = ny * nx, ny_g * nx
    enu = enu_const()
    stream = torch.cuda.Stream()          # the handle launches on this stream, and the events below are recorded on it
    torch.cuda.set_stream(stream)

    cube = DeviceCube.from_dict(cfg['cube'], device=local)
    cube.h.set_stream(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device='cuda')  # > 126 MB L2
    # N > 1: the full maps live in peer-mapped symmetric memory and K3 stores every ray into all GPUs' copies (the all-gather of
    # the output map fused into the integration kernel); NCCL all-gather only when symmetric memory cannot be set up
    sym = comm.symmetric_maps(1, ny_g, nx) if comm else None
    if sym is not None:
        out_w, out_h = sym.maps[0][0, r0:r1], sym.maps[1][0, r0:r1]
        cube.set_exchange(rank, world, sym.xchg_ptrs)   # K0's maxima / counters meet on the device (k_publish -> barrier -> k_plan)
    else:
        out_w = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
        out_h = torch.empty((ny, nx), dtype=torch.float64, device='cuda')
    # host reduce hooks: the fallback when symmetric memory is unavailable, and the route of the row-tiled walk (rasters whose
    # along-ray distances exceed the HBM budget per GPU); the fused step ignores them
    rmax = comm.reduce_max if comm else None
    rsum = comm.reduce_sum if comm else None

    def step():
        if sym is not None:
            # one fused step: K0 -> publish -> barrier -> device plan -> K3 (+ peer stores of its rows) -> barrier; no host in between
            info = cube.trace(_lib.GEOM_GRID, cfg['xpts'], ypts, ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'], cfg['max_segment_length'],
                              out_w, out_h, reduce_max=rmax, reduce_sum=rsum, peers_fn=lambda a, b: sym.peer_ptrs(0, r0 + a, r0 + b), exchange=sym)
            return info, sym.maps[0][0], sym.maps[1][0]
        info = cube.trace(_lib.GEOM_GRID, cfg['xpts'], ypts, ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'], cfg['max_segment_length'],
                          out_w, out_h, reduce_max=rmax, reduce_sum=rsum)
        if comm:
            full_w = comm.all_gather_rows(out_w, ny_g)
            full_h = comm.all_gather_rows(out_h, ny_g)
            return info, full_w, full_h
        return info, out_w, out_h

    def sync_all():
        torch.cuda.synchronize()
        if comm:
            comm.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        info, _, _ = step()
    sync_all()

    # ---- timed region: device-resident value -------------------------------------------------------------------
    l0 = cube.h.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks = ClockSampler(local)
    clocks.__enter__()               # closed after the last timed loop (K2 / K3 kernel timings on rank 0, e2e on the others)
    sync_all()
    for a, b in ev:
        flush.zero_()            # L2 flush between timed iterations (the t-buffer alone is > L2, this makes it explicit)
        a.record(stream)
        info, fw, fh = step()
        b.record(stream)
    sync_all()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    launches = cube.h.launches - l0
    t_local = float(ms.sum())
    if comm:
        t = torch.tensor([t_local], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_local = float(t.item())
    ms_per_step = t_local / args.steps
    value = n_global / (ms_per_step * 1e-3)
    checksum = float(fw.sum().item() + fh.sum().item())
    nan_count = int(torch.isnan(fw).sum().item())
    fused_vs_allgather = sharded_vs_unsharded = None
    if sym is not None:
        # the maps K3 assembled by peer stores against an NCCL all-gather of the same row blocks (outside the timed region)
        fused_vs_allgather = max(float((comm.all_gather_rows(out_w.clone(), ny_g) - fw).abs().max().item()),
                                 float((comm.all_gather_rows(out_h.clone(), ny_g) - fh).abs().max().item()))
        fw, fh = fw.clone(), fh.clone()   # the e2e leg below reuses the symmetric buffers
        # shard == whole: rank 0 retraces 64-row crops (one inside every rank's block) UNSHARDED -- one process, no exchange, the
        # global per-layer maxima handed in -- and compares with the rows the ranks wrote into its maps
        if rank == 0:
            solo = DeviceCube.from_dict(cfg['cube'], device=local)
            solo.h.set_stream(stream.cuda_stream)
            cw = torch.empty((64, nx), dtype=torch.float64, device='cuda')
            ch = torch.empty_like(cw)
            sharded_vs_unsharded = 0.0
            for q in range(world):
                q0, q1 = shard_rows(ny_g, q, world)
                c0 = q0 + max(0, (q1 - q0) // 2 - 32)
                rows = np.ascontiguousarray(cfg['ypts'][c0:c0 + 64])
                solo.ray_layers(_lib.GEOM_GRID, cfg['xpts'], rows, 64, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
                npts, _ = solo.ray_integrate(info.maxlen, cfg['max_segment_length'], info.clamp_low_first, cw, ch)
                assert np.array_equal(npts, info.nparts)
                sharded_vs_unsharded = max(sharded_vs_unsharded, float((cw - fw[c0:c0 + 64]).abs().max().item()),
                                           float((ch - fh[c0:c0 + 64]).abs().max().item()))
            del solo, cw, ch

    # ---- e2e through the public API with host buffers ----------------------------------------------------------
    los = Raytracing(incidence=INC, heading=HEAD)
    cube_host = {k: (torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() if k in ('wet', 'hydro') else v)
                 for k, v in cfg['cube'].items()}
    e2e_steps = max(3, min(args.steps, 10)) if not c5 else 3
    host_block = os.environ.get('RDR_BENCH_E2E_HOSTBLOCK', '1') != '0'

    def e2e_step():
        ifs = getInterpolators(cube_host, device=local)                                   # cube + axes H2D
        if comm:
            from raider_b200.dist import build_cube_ray_sharded
            # every rank's rows go to its own page-locked host block AND into every GPU's full maps, from the same kernel
            res = build_cube_ray_sharded(cfg['xpts'], cfg['ypts'], np.array([0.0]), los, 4326, 4326, list(ifs), comm,
                                         MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'],
                                         gather='device', host_block=host_block)
            if host_block:
                return res[1], res[0]
            rows = [_lib.pinned_empty((1, ny, nx)) for _ in range(2)]     # own rows: one device-to-host copy out of the full maps
            for f in range(2):
                torch.from_numpy(rows[f]).copy_(res[f][:, r0:r1], non_blocking=True)
            torch.cuda.synchronize()
            return rows, res
        return _build_cube_ray(cfg['xpts'], ypts, np.array([0.0]), los, 4326, 4326, list(ifs),   # axes H2D, delay maps D2H
                               MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])

    for _ in range(2):
        res = e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = e2e_step()
    sync_all()
    t_e2e = time.perf_counter() - t0
    if comm:
        t = torch.tensor([t_e2e], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = n_global * e2e_steps / t_e2e
    # What bounds e2e across ranks: all GPUs of the box push their row blocks into host memory at the same time.  Measured here:
    # the plain device-to-host copy of this rank's two row blocks (2 x 8 B per ray, page-locked target) with every rank copying
    # at once -- no kernels, no API around it.  e2e cannot beat ms_per_step (device) + this.
    d2h_ms = None
    if comm and not c5:
        hb = [_lib.pinned_empty((ny, nx)) for _ in range(2)]
        src = [out_w.clone(), out_h.clone()]
        sync_all()
        t0 = time.perf_counter()
        for _ in range(5):
            for f in range(2):
                torch.from_numpy(hb[f]).copy_(src[f], non_blocking=True)
            torch.cuda.synchronize()
        td = torch.tensor([(time.perf_counter() - t0) / 5], dtype=torch.float64, device='cuda')
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
        d2h_ms = 1e3 * float(td.item())
        del hb, src
    cube_bytes = int(cfg['cube']['wet'].nbytes + cfg['cube']['hydro'].nbytes)
    h2d = cube_bytes + 8 * (cfg['xpts'].size + ny + cfg['cube']['x'].size + cfg['cube']['y'].size + cfg['cube']['z'].size)
    d2h = 2 * 8 * n_local
    if comm:   # host rows of this rank vs the device path, and the reassembled map in this GPU's HBM vs the device path's
        e2e_dev_diff = max(float(np.abs(res[0][0][0] - fw[r0:r1].cpu().numpy()).max()), float((res[1][0][0] - fw).abs().max().item()))
    else:
        e2e_dev_diff = float(np.abs(res[0][0] - out_w.cpu().numpy()).max())
    # the same through caller-supplied pageable outputArrs (what RAiDER.delay does when it passes outputArrs, delay.py:245-248,323)
    e2e_pageable = None
    if not comm and not c5:
        outs = [np.zeros((1, ny, nx)), np.zeros((1, ny, nx))]

        def e2e_pageable_step():
            ifs = getInterpolators(cube_host, device=local)
            outs[0][...] = 0.0
            outs[1][...] = 0.0
            _build_cube_ray(cfg['xpts'], ypts, np.array([0.0]), los, 4326, 4326, list(ifs), outputArrs=outs,
                            MAX_SEGMENT_LENGTH=cfg['max_segment_length'], MAX_TROPO_HEIGHT=cfg['zref'])
        e2e_pageable_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_pageable_step()
        torch.cuda.synchronize()
        tp = (time.perf_counter() - t0) / 3
        e2e_pageable = {'value': n_global / tp, 'ms_per_step': 1e3 * tp,
                        'api': '_build_cube_ray(..., outputArrs=<pageable np.zeros arrays>): in-place += on the host, zero-fill of the arrays included',
                        'max_abs_diff_vs_device_path_m': float(np.abs(outs[0][0] - out_w.cpu().numpy()).max())}

    if rank != 0:
        clocks.__exit__()
        if comm:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    roofline = fused = cpu = variants = None
    check = {'checksum': checksum, 'nan': nan_count, 'nparts_sum': int(info.samples_per_ray), 'fused_gather_vs_nccl_allgather_max_abs_diff_m': fused_vs_allgather,
             'sharded_vs_unsharded_max_abs_diff_m': sharded_vs_unsharded, 'knife_edge_redo': bool(info.knife_edge_redo), 'clamp_reruns': int(info.reruns)}
    if not c5:
        # ---- roofline of the unfused trilinear-sample kernel K2 (rank 0, N = 1 semantics) ---------------------------
        nslots = 48
        solo = cube if not comm else DeviceCube.from_dict(cfg['cube'], device=local)
        if comm:
            solo.h.set_stream(stream.cuda_stream)
        solo.ray_layers(_lib.GEOM_GRID, cfg['xpts'], ypts, ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
        maxlen = info.maxlen
        pts = torch.empty((nslots, n_local, 3), dtype=torch.float64, device='cuda')
        solo.ray_points(maxlen, cfg['max_segment_length'], slot0=100, nslots=nslots, out=pts)
        npts = nslots * n_local
        sw = torch.empty(npts, dtype=torch.float64, device='cuda')
        sh = torch.empty(npts, dtype=torch.float64, device='cuda')
        for _ in range(3):
            solo.sample(pts.view(-1, 3), out=(sw, sh))
        torch.cuda.synchronize()
        k2 = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            solo.sample(pts.view(-1, 3), out=(sw, sh))   # 10.2 GB of points + outputs per launch >> L2
            b.record(stream)
            torch.cuda.synchronize()
            k2.append(a.elapsed_time(b))
        k2_ms = float(np.mean(k2))
        k2_bytes = npts * 40 + cube_bytes
        k2_gbs = k2_bytes / (k2_ms * 1e-3) / 1e9
        # DRAM traffic of that launch from the committed ncu --set full capture (dram__bytes_read.sum + dram__bytes_write.sum); only
        # trusted when the capture was taken from the kernel sources of this tree
        k2_traffic, k2_traffic_src = None, None
        tp = ROOT / 'profiles' / 'k2_traffic.json'
        if tp.exists():
            try:
                tj = json.loads(tp.read_text())
                if int(tj['points_per_launch']) == npts and tj.get('kernel_source_hash') == kernel_source_hash():
                    k2_traffic, k2_traffic_src = float(tj['dram_bytes_read'] + tj['dram_bytes_write']), tj['source']
                else:
                    k2_traffic_src = 'profiles/k2_traffic.json was captured from other kernel sources / another launch size: not used'
            except Exception:
                pass
        # fp32-I/O tier of the same kernel (20 B/point)
        pts32 = pts.to(torch.float32)
        sw32 = torch.empty(npts, dtype=torch.float32, device='cuda')
        sh32 = torch.empty(npts, dtype=torch.float32, device='cuda')
        for _ in range(3):
            solo.sample(pts32.view(-1, 3), out=(sw32, sh32))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(10):
            solo.sample(pts32.view(-1, 3), out=(sw32, sh32))
        b.record(stream)
        torch.cuda.synchronize()
        k2_32_gbs = (npts * 20 + cube_bytes) / (a.elapsed_time(b) / 10 * 1e-3) / 1e9
        # CPU samplers beside K2: the installed scipy RGI the reference's delay path calls (delayFcns.py:55-56), 1
        # thread, and the reference's own native RAiDER.interpolate.interpolate compiled from /root/reference (oracle/_ref), with
        # its max_threads = 8 cap (module.cpp:81,293) -- both fields, on a bounded sample of the same points
        cpu_sampler = None
        if world == 1:
            try:
                from scipy.interpolate import RegularGridInterpolator as RGI
                from oracle import build_ref
                n_cpu = 2_000_000
                hp = pts.view(-1, 3)[:: max(1, npts // n_cpu)][:n_cpu].cpu().numpy()
                ys_, xs_, zs_ = (np.asarray(cfg['cube'][k], dtype=np.float64) for k in ('y', 'x', 'z'))
                vals = [np.ascontiguousarray(cfg['cube'][k].transpose(1, 2, 0), dtype=np.float64) for k in ('wet', 'hydro')]
                t0 = time.perf_counter()
                ref_vals = [RGI((ys_, xs_, zs_), v, method='linear', bounds_error=False, fill_value=np.nan)(hp) for v in vals]
                t_scipy = time.perf_counter() - t0
                cpu_sampler = {'sample_points': int(hp.shape[0]), 'scipy_rgi_points_per_s': hp.shape[0] / t_scipy, 'scipy_threads': 1}
                gw = torch.empty(hp.shape[0], dtype=torch.float64, device='cuda')
                gh = torch.empty_like(gw)
                solo.sample(torch.from_numpy(hp).cuda(), out=(gw, gh))
                torch.cuda.synchronize()
                cpu_sampler['k2_bit_identical_to_scipy_on_sample'] = bool(np.array_equal(gw.cpu().numpy(), ref_vals[0], equal_nan=True) and
                                                                         np.array_equal(gh.cpu().numpy(), ref_vals[1], equal_nan=True))
                if build_ref.available():
                    interp = build_ref.load('interpolate')
                    t0 = time.perf_counter()
                    nat = [interp.interpolate((ys_, xs_, zs_), v, hp, fill_value=np.nan, max_threads=8) for v in vals]
                    t_nat = time.perf_counter() - t0
                    cpu_sampler.update({'raider_interpolate_points_per_s': hp.shape[0] / t_nat, 'raider_interpolate_threads': 8,
                                        'raider_interpolate_max_abs_diff_vs_scipy': float(np.nanmax(np.abs(nat[0] - ref_vals[0])))})
            except Exception as e:  # the checker is optional equipment of the bench, never of the product
                cpu_sampler = {'unavailable': repr(e)}
        del pts, pts32, sw, sh, sw32, sh32
        roofline = {'kernel': 'k_sample_stream<double> (K2 trilinear_sample, unfused, TMA-bulk point stream)', 'bound': 'hbm', 'achieved': k2_gbs,
                    'peak': peak, 'unit': 'GB/s', 'frac': k2_gbs / peak, 'traffic': k2_traffic, 'traffic_source': k2_traffic_src,
                    'algorithmic_bytes_per_launch': k2_bytes, 'peak_source': peak_src, 'bytes_per_point': 40,
                    'points_per_launch': npts, 'ms_per_launch': k2_ms, 'fp32_io_tier_gbs': k2_32_gbs, 'fp32_io_tier_frac': k2_32_gbs / peak,
                    'fp32_tier_kernel': 'k_sample_stream_f32 (fp32 coordinates, arithmetic and values, 20 B/point; L1-bandwidth bound: 100 B/point through L1)',
                    'points_per_s': npts / (k2_ms * 1e-3), 'cpu_samplers': cpu_sampler}

        # ---- the two halves of a step alone (events around K0 and around plan + K3) for the fused accounts -----------
        targs = (_lib.GEOM_GRID, cfg['xpts'], ypts, ny, nx, _lib.LOS_ENU_CONST, enu, 0.0, cfg['zref'])
        k0_ms = _events(torch, stream, flush, lambda: solo.trace_begin(*targs), 5)
        solo.trace_begin(*targs)
        tw, th = (out_w.clone(), out_h.clone()) if comm else (out_w, out_h)
        k3_ms = _events(torch, stream, flush, lambda: solo.trace_finish(cfg['max_segment_length'], tw, th), 5)
        uniq = int(info.samples_per_ray - info.n_layers + 1)
        fused = {
            'kernel': 'k_ray_integrate_poly<double> (span cubics + layer quadrature) [+ k_ray_integrate_thin for runs of <= 3-sample layers]; '
                      'flagged rays: k_ray_integrate list pass; plan: k_plan on the device', 'ms': k3_ms, 'k0_ms': k0_ms,
            'bound': 'fp64 issue, three-register DFMAs at 3 cycles (not HBM): see DESIGN.md section 4',
            'algorithmic_bytes_per_ray': 16 + 8 * (info.n_layers + 1),
            'hbm_gbs': n_local * (16 + 8 * (info.n_layers + 1)) / (k3_ms * 1e-3) / 1e9,
            'equivalent_unfused_gbs': n_local * info.samples_per_ray * 40 / (k3_ms * 1e-3) / 1e9,
            'samples_per_ray_reference': info.samples_per_ray, 'unique_samples_per_ray': uniq,
            'samples_per_s': n_local * info.samples_per_ray / (k3_ms * 1e-3),
        }

        # ---- cpu_baseline (N = 1 only): the oracle port on a bounded sample of the same inputs.  nParts comes from the ORACLE'S OWN
        # maxima of the full raster (border + interior sample of losreader.build_ray's restatement) -- nothing borrowed from the GPU
        if world == 1:
            own_max, border_max = cpu_global_maxlen(cfg)
            own_np = np.ceil(own_max / cfg['max_segment_length']).astype(int) + 1
            cpu, want, sl = cpu_baseline_single(cfg, own_max, rows=100, want_out=True)
            got_w, got_h = out_w[sl].cpu().numpy(), out_h[sl].cpu().numpy()
            cpu['max_abs_diff_vs_gpu_m'] = float(max(np.abs(got_w - want[0][0]).max(), np.abs(got_h - want[1][0]).max()))
            cpu['rows_compared'] = 100
            check.update({'nparts_equal_oracle_own_maxima': bool(np.array_equal(own_np, info.nparts)),
                          'maxlen_max_abs_diff_vs_oracle_m': float(np.abs(own_max - info.maxlen).max()),
                          'oracle_max_on_border': bool(np.array_equal(own_max, border_max))})

            # ---- the production shapes beside the headline ("also run the 145-node table variant") -------
            variants = {}
            c145 = syn.config_c2(n=N_SIDE, table='ml145')
            variants['ml145_1000m'] = run_variant(torch, stream, flush, 'C2 raster through the 145-node model-level table (models/model_levels.py:12 shape), 1000 m '
                                                  'segments (the reference default): the production setting of ERA5 / GMAO / HRES cubes', c145, enu)
            c3 = syn.config_c3(ny=N_SIDE, nx=N_SIDE, table='hrrr57')
            variants['hrrr57_lcc'] = run_variant(torch, stream, flush, '2000x2000 rays over a 3 km spherical-Lambert cube (models/hrrr.py:255-260) with the 57-node '
                                                 'table, 1000 m segments, fixed 30 deg incidence', c3, enu, model_crs=c3['crs'])
    clocks.__exit__()

    line = {
        'metric': metric, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong' if c5 else 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {**workload_config(world, cfg['cube'], 'c5' if c5 else 'c2'), 'samples_per_ray': info.samples_per_ray, 'layers': info.n_layers,
                   'l2': 'flushed with a 256 MB write between timed steps', 'parallelism': f'row-block x{world}' if world > 1 else 'single GPU',
                   'row_tiles_per_gpu': info.tiles, 'layers_in_thin_kernel': info.k_split,
                   'output_map_replication': ('NVLink-SHARP multicast: one multimem.st per value from K3' if (sym is not None and sym.mc_base) else
                                              'one peer store per value and GPU from K3' if sym is not None else None),
                   'collectives': ('no NCCL on the data path: K + 3 words per rank stored into every peer\'s exchange slots (symmetric memory), MAX / SUM taken by a one-CTA '
                                   'kernel on every GPU, output maps reassembled on every GPU by peer stores from K3 over NVLink, 2 signal-pad barriers per step'
                                   if sym is not None else 'all_gather_into_tensor (symmetric memory unavailable)') if world > 1 else 'none'},
        'clocks': clocks.summary(),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
                'ms_per_step': 1e3 * t_e2e / e2e_steps,
                'api': ('getInterpolators(host cube) + _build_cube_ray(host axes) -> host float64 maps' if world == 1 else
                        'getInterpolators(host cube) + build_cube_ray_sharded(host axes, gather=device, host_block=' + str(host_block) + '): every rank gets its own '
                        'rows as host float64 arrays and the full maps in HBM; h2d/d2h bytes are per rank'),
                'max_abs_diff_vs_device_path_m': e2e_dev_diff, 'caller_supplied_pageable_outputs': e2e_pageable,
                'concurrent_d2h_of_the_row_blocks_alone_ms': d2h_ms,
                'concurrent_d2h_gbs_per_gpu': (d2h / (d2h_ms * 1e-3) / 1e9) if d2h_ms else None},
        'gpu_launches': int(launches),
        'roofline': roofline,
        'fused': fused,
        'cpu_baseline': cpu,
        'variants': variants,
        'check': check,
    }
    print(json.dumps(line))
    if comm:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c2', choices=['c2', 'c5', 'c5-ml145'],
                    help='c2 (default): BASELINE configs[1], 2000x2000 rays per GPU, weak scaling; c5 / c5-ml145: BASELINE configs[4], 2.4e8 rays '
                         'split over the GPUs (strong scaling), NZ = 72 or the 145-node table')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
