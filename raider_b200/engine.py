"""Host driver of the device delay path: a staged cube + the K0 -> (all-reduce) -> K3 sequence.

This is the layer the Python shims (delay.py, delayFcns.py, losreader.py) sit on.  It owns no numerics: every
number comes out of libraider_b200.so.  Reference structure it replaces: the per-height body of
``_build_cube_ray`` (tools/RAiDER/delay.py:256-323) and ``_build_cube`` (:205-214).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import Handle, NoLayersError, f64, is_device, ptr
from .crs import Geographic, parse_crs


@dataclass
class TraceInfo:
    """What one height slice did: the integer contract (nParts) and the whole-raster predicates."""
    ht: float
    n_rays: int = 0
    n_layers: int = 0
    maxlen: np.ndarray | None = None      # global per-layer max ray length (delay.py:283)
    nparts: np.ndarray | None = None      # int steps per layer
    samples_per_ray: int = 0              # sum(nParts): samples the reference evaluates per ray
    clamp_low_first: bool = False         # delay.py:306-307 fired (first sample below min(z) on every pixel)
    clamp_high_last: bool = False         # delay.py:310-311 fired (last sample above max(z) on every pixel)
    reruns: int = 0
    oob_below: int = 0
    oob_above: int = 0
    n_nan_rays: int = 0
    skipped: bool = False                 # no contributing layer at the last output height (delay.py:276-277)
    tiles: int = 1                        # row tiles the raster was walked in (HBM budget of the t-buffer)
    k_split: int = 0                      # layers integrated by the thin-layer kernel
    n_spans: int = 0
    staged_passes: int = 0                # CTA passes of the thin-layer kernel with TMA-staged record columns / without
    unstaged_passes: int = 0
    knife_edge_redo: bool = False         # nParts hung on the last bits of a maximum: the step was redone with the exact K0
    knife_edge_layer: int = -1            # ... and still does with the exact K0 (reported, delay.py:283)


class DeviceCube:
    """A weather-model cube staged in HBM (replaces the pair of scipy RGIs of delayFcns.py:55-56)."""

    def __init__(self, ys, xs, zs, wet, hydro, layout=_lib.LAYOUT_ZYX, crs=None, device: int | None = None) -> None:
        self.h = Handle(device)
        self.crs = parse_crs(crs)
        # The staged fields are float32 (what the reference's files hold for wet / hydro, weatherModel.py:617-619).  Float64
        # fields (wet_total / hydro_total are float64 in the reference's files, weatherModel.py:398-403; scipy RGIs built on
        # float64 values) are staged as hi + lo float32 parts in two cubes -- value == hi + lo to ~2^-48 -- and sampled /
        # integrated twice: interpolation and the trapezoid sum are linear in the values.
        self.lo = None
        if not is_device(wet) and np.asarray(wet).dtype == np.float64:
            w64, h64 = np.asarray(wet, dtype=np.float64), np.asarray(hydro, dtype=np.float64)
            wet, hydro = w64.astype(np.float32), h64.astype(np.float32)
            with np.errstate(invalid='ignore'):
                w_lo = np.where(np.isfinite(w64), w64 - wet.astype(np.float64), 0.0).astype(np.float32)
                h_lo = np.where(np.isfinite(h64), h64 - hydro.astype(np.float64), 0.0).astype(np.float32)
            if np.any(w_lo) or np.any(h_lo):
                self.lo = DeviceCube(ys, xs, zs, w_lo, h_lo, layout=layout, crs=crs, device=device)
        self.h.set_cube(ys, xs, zs, wet, hydro, layout=layout, crs_kind=self.crs.kind, crs_params=self.crs.params())
        # ascending copies, as scipy exposes them through .grid (read at delay.py:239)
        self.grid = tuple(np.sort(f64(a)) for a in (ys, xs, zs))
        self.shape_zyx = (np.size(zs), np.size(ys), np.size(xs))
        self.last_info: list[TraceInfo] = []

    @property
    def device(self) -> int:
        return self.h.device

    @classmethod
    def from_dict(cls, cube: dict, kind: str = 'pointwise', crs=None, device=None) -> 'DeviceCube':
        """``cube`` = {x, y, z, wet, hydro[, wet_total, hydro_total]} with (z, y, x) fields (the processed-file layout)."""
        w = cube['wet_total' if kind == 'total' else 'wet']
        hy = cube['hydro_total' if kind == 'total' else 'hydro']
        return cls(cube['y'], cube['x'], cube['z'], w, hy, layout=_lib.LAYOUT_ZYX, crs=crs if crs is not None else cube.get('crs'),
                   device=device)

    @classmethod
    def from_epochs(cls, cubes: list, weights: list, kind: str = 'pointwise', crs=None, device=None) -> 'DeviceCube':
        """The temporal interpolation of cli/raider.py:792-833 at staging: ``sum_i w_i * cube_i`` per field.

        Scalar weights ('center_time', cli/raider.py:877-888): float32 arithmetic on the device, exactly the reference's xarray
        expression (python float x float32 array stays float32).  Per-voxel weight arrays ('azimuth_time_grid',
        s1_azimuth_timing.py:337-399; float64, shape (z, y, x)): the reference's product is float64 -- formed here in float64 in its
        order (0 + w_0 f_0 + w_1 f_1 + ...) and staged as hi + lo float32 parts, so sampling and ray tracing see the float64 blend."""
        if len(cubes) != len(weights) or not cubes:
            raise ValueError('one weight (scalar or (z, y, x) array) per epoch')
        fw, fh = ('wet_total', 'hydro_total') if kind == 'total' else ('wet', 'hydro')
        c0 = cubes[0]
        crs = crs if crs is not None else c0.get('crs')
        if all(np.ndim(w) == 0 for w in weights):
            if len(cubes) != 2:
                raise ValueError('scalar weights blend exactly two epochs (cli/raider.py:877-888)')
            cube = cls(c0['y'], c0['x'], c0['z'], c0[fw], c0[fh], layout=_lib.LAYOUT_ZYX, crs=crs, device=device)
            cube.blend(cubes[1][fw], cubes[1][fh], float(weights[0]), float(weights[1]))
            return cube
        shape = np.shape(c0[fw])
        wet = hydro = 0
        for c, w in zip(cubes, weights):
            w = np.broadcast_to(np.asarray(w, dtype=np.float64), shape)
            wet = wet + w * np.asarray(c[fw])
            hydro = hydro + w * np.asarray(c[fh])
        return cls(c0['y'], c0['x'], c0['z'], np.asarray(wet, dtype=np.float64), np.asarray(hydro, dtype=np.float64), layout=_lib.LAYOUT_ZYX,
                   crs=crs, device=device)

    def blend(self, wet1, hydro1, w0: float, w1: float, layout=_lib.LAYOUT_ZYX) -> None:
        """Fused two-epoch temporal interpolation (cli/raider.py:817-819) at staging time."""
        self.h.blend_cube(wet1, hydro1, w0, w1, layout)

    # ------------------------------------------------------------------------------------ K2
    def sample(self, pts, semantics=_lib.SEM_SCIPY, out=None):
        """Both fields at ``pts[..., 3]`` = (y, x, z).  numpy in -> numpy out; torch CUDA in -> torch CUDA out."""
        if is_device(pts):
            import torch
            dt = _lib.F32 if pts.dtype == torch.float32 else _lib.F64
            if dt == _lib.F64 and pts.dtype != torch.float64:
                raise TypeError('device points must be float32 or float64')
            p = pts.contiguous()
            n = p.numel() // 3
            if out is None:
                out = (torch.empty(p.shape[:-1], dtype=p.dtype, device=p.device), torch.empty(p.shape[:-1], dtype=p.dtype, device=p.device))
            self.h.call('rdr_sample', ptr(p), n, ptr(out[0]), ptr(out[1]), dt, semantics, _lib.MEM_DEVICE)
            if self.lo is not None:
                lw, lh = self.lo.sample(pts, semantics)
                out[0].add_(lw)
                out[1].add_(lh)
            return out
        p = np.asarray(pts)
        if p.dtype != np.float32:
            p = p.astype(np.float64, copy=False)
        p = np.ascontiguousarray(p)
        if p.shape[-1] != 3:
            raise ValueError(f'The requested sample points xi have dimension {p.shape[-1]} but this interpolator has dimension 3')
        n = p.size // 3
        w = np.empty(p.shape[:-1], dtype=p.dtype)
        hy = np.empty(p.shape[:-1], dtype=p.dtype)
        self.h.call('rdr_sample', ptr(p), n, ptr(w), ptr(hy), _lib.F32 if p.dtype == np.float32 else _lib.F64, semantics, _lib.MEM_HOST)
        if self.lo is not None:
            lw, lh = self.lo.sample(p, semantics)
            w += lw
            hy += lh
        return w, hy

    def sample_grid(self, xpts, ypts, ht: float):
        """One height of ``_build_cube`` (delay.py:205-214) when the query grid is in the cube's own CRS."""
        xpts, ypts = f64(xpts), f64(ypts)
        w = np.empty((ypts.size, xpts.size))
        hy = np.empty((ypts.size, xpts.size))
        self.h.call('rdr_sample_grid', ptr(xpts), xpts.size, ptr(ypts), ypts.size, float(ht), ptr(w), ptr(hy), _lib.MEM_HOST)
        if self.lo is not None:
            lw, lh = self.lo.sample_grid(xpts, ypts, ht)
            w += lw
            hy += lh
        return w, hy

    def sample_grid_levels(self, xpts, ypts, zpts):
        """Every height of ``_build_cube`` (delay.py:205-214) in one launch: two (nh, ny, nx) arrays."""
        xpts, ypts, zpts = f64(xpts), f64(ypts), f64(np.atleast_1d(zpts))
        w = np.empty((zpts.size, ypts.size, xpts.size))
        hy = np.empty((zpts.size, ypts.size, xpts.size))
        self.h.call('rdr_sample_grid_levels', ptr(xpts), xpts.size, ptr(ypts), ypts.size, ptr(zpts), zpts.size, ptr(w), ptr(hy), _lib.MEM_HOST)
        if self.lo is not None:
            lw, lh = self.lo.sample_grid_levels(xpts, ypts, zpts)
            w += lw
            hy += lh
        return w, hy

    # ------------------------------------------------------------------------------------ K0 + K3
    def ray_plan(self, ht: float, zref: float):
        n = C.c_int64(0)
        nz = self.grid[2].size
        lo, hi = np.empty(nz), np.empty(nz)
        self.h.call('rdr_ray_plan', float(ht), float(zref), C.byref(n), ptr(lo), ptr(hi))
        return lo[: n.value].copy(), hi[: n.value].copy()

    def ray_layers(self, geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref):
        """K0.  Returns (maxlen[K], counts[5] = {rays, NaN rays, first samples below min(z), K, last samples above max(z)}) for this
        device's rays; raises NoLayersError when K == 0."""
        nz = self.grid[2].size
        maxlen = np.zeros(nz)
        counts = np.zeros(5, dtype=np.int64)
        dev = is_device(gx) or is_device(los)
        self._keep_geom = (gx, gy, los)  # device pointers must outlive rdr_ray_integrate
        self._n_rays = int(ny) * int(nx)
        self.h.call('rdr_ray_layers', geom_kind, ptr(gx), ptr(gy), int(ny), int(nx), los_kind, ptr(los), float(ht), float(zref),
                    ptr(maxlen), ptr(counts), _lib.MEM_DEVICE if dev else _lib.MEM_HOST)
        return maxlen[: counts[3]].copy(), counts

    def ray_integrate(self, maxlen, max_segment_length, clamp_low_first, out_wet, out_hydro, accumulate=False, peers=None):
        """K3 on the rays of the last ray_layers call.  Returns (nparts[K], oob[3]).

        ``peers = (wet_ptrs, hydro_ptrs)``: device addresses (ints) of the same row block inside the delay maps of the node's
        other GPUs (peer-mapped symmetric memory); the kernel mirrors every ray's results there as it finishes the ray
        (rdr_set_peer_outputs) -- the all-gather of the output map fused into the integration.
        """
        if peers is not None and len(peers) == 3 and peers[2] == 'multicast':
            self.h.call('rdr_set_multicast_outputs', int(peers[0][0]), int(peers[1][0]))
            try:
                return self.ray_integrate(maxlen, max_segment_length, clamp_low_first, out_wet, out_hydro, accumulate)
            finally:
                self.h.call('rdr_set_multicast_outputs', None, None)
        if peers is not None and len(peers[0]):
            n = len(peers[0])
            pw, ph = (C.c_void_p * n)(*[int(a) for a in peers[0]]), (C.c_void_p * n)(*[int(a) for a in peers[1]])
            self.h.call('rdr_set_peer_outputs', n, pw, ph)
            try:
                return self.ray_integrate(maxlen, max_segment_length, clamp_low_first, out_wet, out_hydro, accumulate)
            finally:
                self.h.call('rdr_set_peer_outputs', 0, None, None)
        maxlen = f64(maxlen)
        nparts = np.zeros(maxlen.size, dtype=np.int64)
        oob = np.zeros(3, dtype=np.int64)
        dev = is_device(out_wet)
        if dev:
            import torch
            dt = _lib.F32 if out_wet.dtype == torch.float32 else _lib.F64
        else:
            dt = _lib.F32 if out_wet.dtype == np.float32 else _lib.F64
        # (clamp_low_first: bool, or the two-bit flag word of the ABI -- bit 0 lower clamp of the first sample, bit 1 upper clamp of the last)
        self.h.call('rdr_ray_integrate', ptr(maxlen), float(max_segment_length), int(clamp_low_first) & 3, ptr(out_wet), ptr(out_hydro),
                    dt, int(bool(accumulate)), ptr(nparts), ptr(oob), _lib.MEM_DEVICE if dev else _lib.MEM_HOST)
        return nparts, oob

    def ray_points(self, maxlen, max_segment_length, slot0=0, nslots=None, out=None, dtype=np.float64):
        """K1b: sample points (y, x, z) of the last ray_layers call, shape (nslots, n_rays, 3); out may be a torch CUDA tensor."""
        maxlen = f64(maxlen)
        total = C.c_int64(0)
        self.h.call('rdr_ray_points', ptr(maxlen), float(max_segment_length), 0, 0, None, _lib.F64, C.byref(total), _lib.MEM_HOST)
        if nslots is None:
            nslots = total.value - slot0
        if out is None:
            n_rays = int(self._n_rays)
            out = np.empty((nslots, n_rays, 3), dtype=dtype)
        dev = is_device(out)
        if dev:
            import torch
            dt = _lib.F32 if out.dtype == torch.float32 else _lib.F64
        else:
            dt = _lib.F32 if out.dtype == np.float32 else _lib.F64
        self.h.call('rdr_ray_points', ptr(maxlen), float(max_segment_length), int(slot0), int(nslots), ptr(out), dt, C.byref(total),
                    _lib.MEM_DEVICE if dev else _lib.MEM_HOST)
        return out, total.value

    def trace_stations(self, lon, lat, hgt, los_kind, los, zref, max_segment_length):
        """K5: every point is its own 1 x 1 raster at its own height (BASELINE C4, GNSS stations).

        Returns ``(wet[n], hydro[n], nsamples[n])``; numpy in -> numpy out, torch CUDA in -> torch CUDA out.
        """
        dev = is_device(lon)
        if dev:
            import torch
            n = lon.numel()
            wet = torch.empty(n, dtype=torch.float64, device=lon.device)
            hydro = torch.empty_like(wet)
            ns = torch.empty(n, dtype=torch.int32, device=lon.device)
        else:
            lon, lat, hgt = f64(lon).ravel(), f64(lat).ravel(), f64(hgt).ravel()
            n = lon.size
            wet, hydro, ns = np.empty(n), np.empty(n), np.empty(n, dtype=np.int32)
            if los is not None:
                los = f64(los)
        self.h.call('rdr_ray_stations', ptr(lon), ptr(lat), ptr(hgt), int(n), int(los_kind), ptr(los), float(zref), float(max_segment_length),
                    ptr(wet), ptr(hydro), ptr(ns), _lib.MEM_DEVICE if dev else _lib.MEM_HOST)
        return wet, hydro, ns

    def trace(self, geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, max_segment_length, out_wet, out_hydro,
              reduce_max=None, reduce_sum=None, max_t_bytes=None, peers_fn=None, exchange=None) -> TraceInfo:
        """One output height: K0 -> global reduction of the per-layer maxima / predicates -> K3.

        ``reduce_max`` / ``reduce_sum`` are the cross-GPU hooks (numpy array in -> reduced numpy array out); they are
        what keeps ``nParts`` (delay.py:283) and the ``.all()`` clamp (delay.py:306-307) *global* when the raster is
        sharded over ranks.  Outputs are written (not accumulated) into out_wet/out_hydro.

        The along-ray distances K0 hands to K3 take 8 (K+1) bytes per ray.  When that exceeds ``max_t_bytes`` (default
        RAIDER_B200_T_BUDGET_GB, 64 GB of the 180 GB HBM3e) the raster is walked in row tiles: a first K0 pass over all tiles
        for the global maxima and counters, then K0 + K3 per tile with those maxima -- the same mechanism that keeps
        multi-GPU shards identical to the unsharded raster.

        ``peers_fn(r0, r1)`` -> ``(wet_ptrs, hydro_ptrs)`` for rows [r0, r1) of this call's block: see ``ray_integrate``.
        """
        if self.lo is not None:
            # float64 refractivity fields (the reference's files hold float32, weatherModel.py:617-619): integrate the hi and the lo
            # parts and add -- the trapezoid sums are linear in the values; nParts and the predicates are geometry only
            lo, self.lo = self.lo, None
            try:
                info = self.trace(geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, max_segment_length, out_wet, out_hydro,
                                  reduce_max, reduce_sum, max_t_bytes, peers_fn, exchange)
            finally:
                self.lo = lo
            if is_device(out_wet):
                import torch
                tw, th = torch.empty_like(out_wet), torch.empty_like(out_hydro)
            else:
                tw, th = _lib.pinned_empty(out_wet.shape, out_wet.dtype), _lib.pinned_empty(out_hydro.shape, out_hydro.dtype)
            lo.trace(geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, max_segment_length, tw, th, reduce_max, reduce_sum, max_t_bytes,
                     None, exchange)
            out_wet += tw
            out_hydro += th
            if peers_fn is not None:
                raise NotImplementedError('peer-mirrored outputs with float64 refractivity fields')
            return info
        if max_t_bytes is None:
            max_t_bytes = float(os.environ.get('RAIDER_B200_T_BUDGET_GB', '64')) * 2**30
        nz = self.grid[2].size
        rows_per_tile = int(max(1, max_t_bytes // (8 * nz * max(1, int(nx)))))
        if rows_per_tile >= ny:
            return self._trace_block(geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, max_segment_length, out_wet, out_hydro,
                                     reduce_max, reduce_sum, peers_fn(0, ny) if peers_fn else None, exchange)
        # (the row-tiled walk reduces through the host hooks even when a device exchange is attached: its first pass visits every
        # tile before any plan can be made; the exchange only closes the step with its barrier)
        tiles = [(r0, min(ny, r0 + rows_per_tile)) for r0 in range(0, ny, rows_per_tile)]

        def block(r0, r1):
            if geom_kind == _lib.GEOM_GRID:
                return gx, gy[r0:r1], (los if los_kind != _lib.LOS_ARRAY else los[r0 * nx:r1 * nx])
            return gx[r0 * nx:r1 * nx], gy[r0 * nx:r1 * nx], (los if los_kind != _lib.LOS_ARRAY else los[r0 * nx:r1 * nx])

        # pass 1: maxima and counters of every tile
        maxlen, counts = None, None
        for r0, r1 in tiles:
            bx, by, bl = block(r0, r1)
            m, c = self.ray_layers(geom_kind, bx, by, r1 - r0, nx, los_kind, bl, ht, zref)
            maxlen = m if maxlen is None else np.maximum(maxlen, m)
            if counts is None:
                counts = c.copy()
            else:
                counts[_SUMMED] += c[_SUMMED]
        maxlen, counts = _global_plan(maxlen, counts, reduce_max, reduce_sum)
        info = TraceInfo(ht=float(ht))
        info.n_rays, info.n_nan_rays, info.n_layers = int(counts[0]), int(counts[1]), int(counts[3])
        if counts[1] == counts[0]:
            raise ValueError('geo2rdr did not converge. Check orbit coverage')  # delay.py:279-280, over the whole raster
        clamp = _clamp_flags(counts)
        ow = out_wet.reshape(ny, nx) if hasattr(out_wet, 'reshape') else out_wet
        oh = out_hydro.reshape(ny, nx) if hasattr(out_hydro, 'reshape') else out_hydro
        for attempt in range(2):
            oob_tot = np.zeros(3, dtype=np.int64)
            for r0, r1 in tiles:
                bx, by, bl = block(r0, r1)
                self.ray_layers(geom_kind, bx, by, r1 - r0, nx, los_kind, bl, ht, zref)
                nparts, oob = self.ray_integrate(maxlen, max_segment_length, clamp, ow[r0:r1], oh[r0:r1],
                                                 peers=peers_fn(r0, r1) if peers_fn else None)
                oob_tot += oob
            # K3 re-evaluates the first-sample predicate on its own heights (bitwise K0's for all but polar / projected-cube rays)
            below3 = oob_tot[:1] if reduce_sum is None else reduce_sum(oob_tot[:1])
            if bool(below3[0] == counts[0]) == bool(clamp & 1):
                break
            clamp ^= 1  # K0's hint and K3's own evaluation disagree on a knife edge: K3 rules
            info.reruns = 1
        info.maxlen, info.nparts = maxlen, nparts
        info.samples_per_ray = int(nparts.sum())
        info.clamp_low_first, info.clamp_high_last = bool(clamp & 1), bool(clamp & 2)
        info.oob_below, info.oob_above = int(oob_tot[1]), int(oob_tot[2])
        info.tiles = len(tiles)
        if exchange is not None:
            exchange.barrier()   # all rows of all ranks are in the maps
        _check_clamp_coverage(info, max_segment_length)
        return info

    # ------------------------------------------------------------------------------------ fused step (device-side plan)
    def trace_begin(self, geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, flags=0):
        """K0 (+ the cross-GPU publication of its K + 3 words when an exchange is attached); only enqueues."""
        dev = is_device(gx) or is_device(los)
        self._keep_geom = (gx, gy, los)  # device pointers must outlive the step
        self._n_rays = int(ny) * int(nx)
        self.h.call('rdr_trace_begin', geom_kind, ptr(gx), ptr(gy), int(ny), int(nx), los_kind, ptr(los), float(ht), float(zref),
                    int(flags), _lib.MEM_DEVICE if dev else _lib.MEM_HOST)

    def trace_finish(self, max_segment_length, out_wet, out_hydro, force_clamp=-1, mode=_lib.K3_AUTO, accumulate=False, peers=None):
        """k_plan + K3 on the device plan; only enqueues.  ``peers`` as in :meth:`ray_integrate`."""
        dev = is_device(out_wet)
        if dev:
            import torch
            dt = _lib.F32 if out_wet.dtype == torch.float32 else _lib.F64
        else:
            dt = _lib.F32 if out_wet.dtype == np.float32 else _lib.F64
        args = ('rdr_trace_finish', float(max_segment_length), int(force_clamp), int(mode), ptr(out_wet), ptr(out_hydro), dt,
                int(bool(accumulate)), _lib.MEM_DEVICE if dev else _lib.MEM_HOST)
        if peers is not None and len(peers) == 3 and peers[2] == 'multicast':
            # ONE multicast destination: the NVSwitch replicates every store into all GPUs' maps (rdr_set_multicast_outputs)
            self.h.call('rdr_set_multicast_outputs', int(peers[0][0]), int(peers[1][0]))
            try:
                self.h.call(*args)
            finally:
                self.h.call('rdr_set_multicast_outputs', None, None)
        elif peers is not None and len(peers[0]):
            n = len(peers[0])
            pw, ph = (C.c_void_p * n)(*[int(a) for a in peers[0]]), (C.c_void_p * n)(*[int(a) for a in peers[1]])
            self.h.call('rdr_set_peer_outputs', n, pw, ph)
            try:
                self.h.call(*args)
            finally:
                self.h.call('rdr_set_peer_outputs', 0, None, None)
        else:
            self.h.call(*args)

    def trace_result(self):
        """Synchronise and read the step plan back: (maxlen[K], nparts[K], info[20]) -- see rdr_trace_result."""
        nz = self.grid[2].size
        maxlen, nparts, info = np.zeros(nz), np.zeros(nz, dtype=np.int64), np.zeros(20, dtype=np.int64)
        self.h.call('rdr_trace_result', ptr(maxlen), ptr(nparts), ptr(info))
        k = int(info[2])
        return maxlen[:k].copy(), nparts[:k].copy(), info

    def set_exchange(self, rank: int, world: int, bufs) -> None:
        """Attach (world > 0) / detach the peer-mapped exchange buffers of the cross-GPU plan (rdr_set_exchange)."""
        if world:
            arr = (C.c_void_p * world)(*[int(a) for a in bufs])
            self.h.call('rdr_set_exchange', int(rank), int(world), arr)
        else:
            self.h.call('rdr_set_exchange', 0, 0, None)

    def _trace_block(self, geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, max_segment_length, out_wet, out_hydro,
                     reduce_max=None, reduce_sum=None, peers=None, exchange=None) -> TraceInfo:
        """One height of one row block as ONE fused step: K0 -> [exchange barrier] -> device plan -> K3, no host in between.

        ``exchange``: an object with ``barrier()`` (signal-pad barrier on the kernels' stream; raider_b200.dist.SymmetricMaps)
        whose buffers were attached with :meth:`set_exchange` -- the per-layer maxima and predicate counters of all ranks then meet
        on the device.  Host hooks (``reduce_max`` / ``reduce_sum`` without an exchange: gloo, no P2P) take the unfused route.
        """
        if reduce_max is not None and exchange is None:
            return self._trace_block_hosted(geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, max_segment_length, out_wet, out_hydro,
                                            reduce_max, reduce_sum, peers)
        info = TraceInfo(ht=float(ht))
        flags, force_clamp, mode = 0, -1, _lib.K3_AUTO
        for attempt in range(4):
            self.trace_begin(geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, flags)
            if exchange is not None:
                exchange.barrier()       # every rank's K0 words have landed in every rank's slots (and the previous maps are free)
            self.trace_finish(max_segment_length, out_wet, out_hydro, force_clamp=force_clamp, mode=mode, peers=peers)
            if exchange is not None:
                exchange.barrier()       # all rows of all ranks are in the maps; K3's cross-check counts have landed
            maxlen, nparts, res = self.trace_result()
            status, blocked = int(res[0]), int(res[1])
            info.n_rays, info.n_nan_rays, info.n_layers = int(res[3]), int(res[4]), int(res[2])
            if blocked & _lib.PLAN_ALL_NAN:
                raise ValueError('geo2rdr did not converge. Check orbit coverage')  # delay.py:279-280, over the whole raster
            if blocked & _lib.PLAN_ABSURD:
                raise ValueError('a per-layer maximum ray length is NaN or absurd: nParts (delay.py:283) is undefined')
            if blocked & _lib.PLAN_KNIFE_EDGE:
                # nParts = ceil(max / S) + 1 hangs on the last bits of a maximum: redo the step with the exact (Bowring) K0
                import logging
                logging.getLogger('raider_b200').warning(
                    'nParts knife edge: max ray length of layer %d is %.9f m = %.9f segments of %g m; redoing the step with the exact K0',
                    int(res[11]), maxlen[int(res[11])], maxlen[int(res[11])] / max_segment_length, max_segment_length)
                flags |= _lib.TRACE_EXACT_K0
                info.knife_edge_redo = True
                continue
            if blocked & _lib.PLAN_SPAN_TOO_LONG:
                mode = _lib.K3_FAST
                continue
            if status & _lib.PLAN_KNIFE_EDGE:
                import logging
                logging.getLogger('raider_b200').warning(
                    'nParts knife edge persists with the exact K0: max ray length of layer %d is within 1e-6 segments of a multiple of %g m; '
                    'the step count of that layer may differ from the reference by one', int(res[11]), max_segment_length)
                info.knife_edge_layer = int(res[11])
            clamp = bool(res[7])
            # K3 re-evaluates the first-sample predicate on its own heights (bitwise K0's for all but polar / projected-cube rays);
            # its global count rode the exchange.  On a knife edge K3 rules: redo with the clamp forced.
            if bool(res[6] == res[3]) != clamp and force_clamp < 0:
                force_clamp = int(not clamp)
                info.reruns += 1
                continue
            break
        info.maxlen, info.nparts = maxlen, nparts
        info.samples_per_ray = int(nparts.sum())
        info.clamp_low_first, info.clamp_high_last = bool(res[7]), bool(res[19])
        info.oob_below, info.oob_above = int(res[8]), int(res[9])
        info.k_split, info.n_spans = int(res[12]), int(res[13])
        info.staged_passes, info.unstaged_passes = int(res[16]), int(res[17])
        _check_clamp_coverage(info, max_segment_length)
        return info

    def _trace_block_hosted(self, geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref, max_segment_length, out_wet, out_hydro,
                            reduce_max=None, reduce_sum=None, peers=None) -> TraceInfo:
        """The unfused route: K0, the reductions through host hooks (any torch.distributed backend), K3."""
        info = TraceInfo(ht=float(ht))
        maxlen, counts = self.ray_layers(geom_kind, gx, gy, ny, nx, los_kind, los, ht, zref)
        maxlen, counts = _global_plan(maxlen, counts, reduce_max, reduce_sum)
        info.n_rays, info.n_nan_rays, info.n_layers = int(counts[0]), int(counts[1]), int(counts[3])
        if counts[1] == counts[0]:
            raise ValueError('geo2rdr did not converge. Check orbit coverage')  # delay.py:279-280, over the whole raster
        clamp = _clamp_flags(counts)
        nparts, oob = self.ray_integrate(maxlen, max_segment_length, clamp, out_wet, out_hydro, peers=peers)
        below3 = oob[:1] if reduce_sum is None else reduce_sum(oob[:1])   # K3's own evaluation of the predicate, summed over ranks
        if bool(below3[0] == counts[0]) != bool(clamp & 1):  # K0's hint and K3's own evaluation disagree on a knife edge: K3 rules
            clamp ^= 1
            nparts, oob = self.ray_integrate(maxlen, max_segment_length, clamp, out_wet, out_hydro, peers=peers)
            info.reruns = 1
        info.maxlen, info.nparts = maxlen, nparts
        info.samples_per_ray = int(nparts.sum())
        info.clamp_low_first, info.clamp_high_last = bool(clamp & 1), bool(clamp & 2)
        info.oob_below, info.oob_above = int(oob[1]), int(oob[2])
        return info


def _check_clamp_coverage(info: 'TraceInfo', max_segment_length: float) -> None:
    """Two of the `.all()` clamps of delay.py:306-311 are applied on the device: the first sample of a ray (all pixels below min(z))
    and the last one (all pixels above max(z): the default zref = top - 1 m with rays steeper than ~58 deg).  The others (an interior
    sample slot with ALL pixels outside the vertical range) cannot fire for valid inputs; should every ray nevertheless have left the
    model vertically at such a slot, the device returned NaN where the reference clamps: say so instead of returning it silently."""
    n = max(info.n_rays, 1)
    if info.oob_above >= n or (info.oob_below >= n and not info.clamp_low_first):
        import logging
        logging.getLogger('raider_b200').critical(
            'every ray has samples outside the vertical range of the model (below: %d, above: %d of %d rays): the reference clamps such '
            'samples to min(z) / max(z) when ALL pixels of a sample slot are outside (delay.py:306-311); the device path returns NaN',
            info.oob_below, info.oob_above, info.n_rays)


def _global_plan(maxlen, counts, reduce_max, reduce_sum):
    """Per-layer maxima (MAX) and the three predicate counters (SUM) over all ranks.  When the hooks belong to an object that
    offers ``reduce_pair`` (raider_b200.dist.Comm) both travel in ONE collective instead of two."""
    if reduce_max is None:
        return maxlen, counts
    pair = getattr(getattr(reduce_max, '__self__', None), 'reduce_pair', None)
    if pair is not None:
        m, c4 = pair(maxlen, counts[_SUMMED])
    else:
        m, c4 = reduce_max(maxlen), reduce_sum(counts[_SUMMED])
    out = counts.copy()
    out[_SUMMED] = c4
    return m, out


_SUMMED = [0, 1, 2, 4]   # of the counts of rdr_ray_layers: rays, NaN rays, first samples below min(z), [3 = K], last samples above max(z)


def _clamp_flags(counts) -> int:
    """The two whole-raster `.all()` predicates of delay.py:306-311 from K0's global counts: bit 0 = every pixel's first sample lies
    below min(z), bit 1 = every pixel's last sample lies above max(z)."""
    return int(counts[2] == counts[0]) | (int(counts[4] == counts[0]) << 1)


def los_device_spec(los, ny: int, nx: int):
    """(los_kind, los_payload) for LOS objects that can be generated on the device, else None."""
    spec = getattr(los, 'device_spec', None)
    if spec is None:
        return None
    return spec()
