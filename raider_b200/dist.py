"""Raster sharding over the GPUs of one node: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch).

The reference has no parallel delay path (``nproc > 1`` raises NotImplementedError, tools/RAiDER/delay.py:178-185).
Rays are independent except for three whole-raster quantities, which is why naive tiling changes results
(SURVEY.md section 0 facts 4-5):

* ``nParts[k] = ceil(max_over_raster(ray_length[k]) / MAX_SEGMENT_LENGTH) + 1``      (delay.py:283)
* the ``.all()`` clamp predicates on the sample heights                              (delay.py:306-311)
* the ``isnan(ray_lengths).all()`` convergence check                                 (delay.py:279)

So the path shards as: contiguous row blocks of the query raster per rank (cube replicated, it is MBs), K0 on each
rank, ONE exchange carrying the K per-layer maxima (MAX) and the 3 predicate counters (SUM) of every rank, K3 on each rank,
and the reassembly of the two output maps so every rank holds the full delay map.  There is no other data-path exchange.
Under NCCL with peer access the exchange never touches the host: every rank's K0 stores its K + 3 words into a slot of every
peer's exchange buffer (symmetric memory, ``rdr_set_exchange``), one signal-pad barrier orders the ranks on the stream, and a
one-CTA kernel on every rank takes MAX / SUM over the slots and builds the step plan on the device (``k_plan``).  K3's own
re-evaluation of the clamp predicate of delay.py:306-307 rides the same slots (one more word per rank) and is cross-checked
after the closing barrier, exactly as in a single-process run.

The reassembly is fused into K3 (``SymmetricMaps``): the full maps live in symmetric memory (``torch.distributed.
_symmetric_memory``: every rank's buffer is peer-mapped into every other rank over NVLink / NVSwitch), and the integration
kernel stores each finished ray into the row block of *every* GPU's maps (``rdr_set_peer_outputs``): 16 B per ray and peer of
posted NVLink writes spread over the whole integration, then one signal-pad barrier -- no all-gather after the kernel.  When
symmetric memory cannot be set up (gloo, no P2P) the same maps are assembled with ``all_gather_into_tensor``.

The same code runs under ``gloo`` on CPU tensors (tests/test_dist_gloo.py drives it with the oracle standing in for the
kernels) and under ``nccl`` on CUDA tensors.
"""
from __future__ import annotations

import numpy as np


def shard_rows(ny: int, rank: int, world: int):
    """Contiguous, near-equal row block [r0, r1) of rank ``rank``; the first ``ny % world`` ranks get one extra row."""
    base, rem = divmod(int(ny), int(world))
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


class Comm:
    """The two reductions and the gather the path needs, on whatever backend the default group uses."""

    def __init__(self, group=None, device=None) -> None:
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        backend = dist.get_backend(group)
        self.device = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
        if device is not None:
            self.device = torch.device(device)

    def _allreduce(self, arr, op):
        t = self.torch.as_tensor(np.ascontiguousarray(arr)).to(self.device)
        self.dist.all_reduce(t, op=op, group=self.group)
        return t.cpu().numpy()

    def reduce_max(self, arr):
        """all-reduce(MAX) of the per-layer maxima: keeps nParts global (delay.py:283)."""
        return self._allreduce(np.asarray(arr, dtype=np.float64), self.dist.ReduceOp.MAX)

    def reduce_sum(self, arr):
        """all-reduce(SUM) of the predicate counters: keeps the .all() tests global (delay.py:279,306-311)."""
        return self._allreduce(np.asarray(arr, dtype=np.int64), self.dist.ReduceOp.SUM)

    def reduce_pair(self, maxlen, counts):
        """Both reductions of the step in one collective: every rank contributes K maxima + 3 counters (as doubles: counts stay
        far below 2^53), one all-gather of world x (K + 3) doubles, MAX / SUM taken locally."""
        torch = self.torch
        maxlen, counts = np.asarray(maxlen, dtype=np.float64), np.asarray(counts, dtype=np.int64)
        mine = torch.as_tensor(np.concatenate([maxlen, counts.astype(np.float64)])).to(self.device)
        flat = torch.empty(self.world * mine.numel(), dtype=torch.float64, device=self.device)  # (gloo wants the flat layout)
        self.dist.all_gather_into_tensor(flat, mine, group=self.group)
        allv = flat.cpu().numpy().reshape(self.world, mine.numel())
        k = maxlen.size
        # NaN maxima (a rank whose rays all failed) must not win or vanish silently: np.max propagates NaN like the all-reduce(MAX) of NCCL does not;
        # the reference's own nanmax-free `.max()` (delay.py:283) propagates, so propagate
        return allv[:, :k].max(axis=0), np.rint(allv[:, k:].sum(axis=0)).astype(np.int64)

    def all_gather_rows(self, block, ny: int):
        """Reassemble an (ny, nx) map from per-rank row blocks (uneven blocks are padded to the largest)."""
        torch = self.torch
        is_t = torch.is_tensor(block)
        t = block if is_t else torch.as_tensor(np.ascontiguousarray(block))
        t = t.to(self.device)
        nx = t.shape[-1]
        if int(ny) % self.world == 0:  # even blocks: one collective straight into the full map, no padding, no reassembly
            full = torch.empty((int(ny), nx), dtype=t.dtype, device=self.device)
            self.dist.all_gather_into_tensor(full, t.contiguous(), group=self.group)
            return full if is_t else full.cpu().numpy()
        rows_max = -(-int(ny) // self.world)
        pad = torch.zeros((rows_max, nx), dtype=t.dtype, device=self.device)
        pad[: t.shape[0]] = t
        out = torch.empty((self.world * rows_max, nx), dtype=t.dtype, device=self.device)
        self.dist.all_gather_into_tensor(out, pad, group=self.group)
        parts = []
        for r in range(self.world):
            r0, r1 = shard_rows(ny, r, self.world)
            parts.append(out[r * rows_max: r * rows_max + (r1 - r0)])
        full = torch.cat(parts, dim=0)
        return full if is_t else full.cpu().numpy()

    def barrier(self):
        self.dist.barrier(group=self.group)

    def symmetric_maps(self, nz: int, ny: int, nx: int):
        """The (2, nz, ny, nx) float64 delay maps of this rank in peer-mapped symmetric memory, or None when that is unavailable.
        One rendezvous per shape; the buffers are reused by later calls of the same shape."""
        cache = self.__dict__.setdefault('_symm', {})
        key = (int(nz), int(ny), int(nx))
        if key not in cache:
            try:
                cache[key] = SymmetricMaps(self, *key)
            except Exception as e:  # no P2P / unsupported backend: NCCL all-gather instead
                import logging
                logging.getLogger(__name__).warning('symmetric memory unavailable (%r): falling back to all_gather_into_tensor', e)
                cache[key] = None
        return cache[key]


class SymmetricMaps:
    """Full (wet, hydro) delay maps replicated in the HBM of every rank, written by peer stores from the integration kernels."""

    def __init__(self, comm: Comm, nz: int, ny: int, nx: int) -> None:
        import torch
        import torch.distributed._symmetric_memory as symm
        if comm.device.type != 'cuda':
            raise RuntimeError('symmetric memory needs the nccl backend on CUDA devices')
        if comm.world > 8:
            raise RuntimeError('rdr_set_peer_outputs takes at most 8 destinations')
        self.comm, self.nz, self.ny, self.nx = comm, nz, ny, nx
        self.maps = symm.empty((2, nz, ny, nx), dtype=torch.float64, device=comm.device)
        group = comm.group if comm.group is not None else comm.dist.group.WORLD
        self.hdl = symm.rendezvous(self.maps, group)
        self.base = [int(a) for a in self.hdl.buffer_ptrs]
        if len(self.base) != comm.world or self.base[comm.rank] != self.maps.data_ptr():
            raise RuntimeError('symmetric-memory rendezvous returned unexpected buffer pointers')
        # NVLink-SHARP multicast mapping of the maps (0 when the fabric / driver offers none): one multimem.st reaches every GPU.
        # Opt-in (RAIDER_B200_MULTICAST=1): measured on 8 x B200 it is correct (maps == NCCL all-gather) but not faster than one
        # plain peer store per GPU -- 3.82 vs 3.65 .. 3.87 ms per C2 step (profiles/r02g_*) -- so the proven path stays the default
        import os
        self.mc_base = int(getattr(self.hdl, 'multicast_ptr', 0) or 0) if os.environ.get('RAIDER_B200_MULTICAST') == '1' else 0
        # exchange slots of the device-side plan (rdr_set_exchange): 2 parities x world slots of K + 3 words per rank
        from . import _lib
        words = int(_lib.load().rdr_exchange_bytes(comm.world)) // 8
        self.xchg = symm.empty((words,), dtype=torch.int64, device=comm.device)
        self.xchg.zero_()
        self.xchg_hdl = symm.rendezvous(self.xchg, group)
        self.xchg_ptrs = [int(a) for a in self.xchg_hdl.buffer_ptrs]
        self.stream = torch.cuda.Stream(device=comm.device)
        torch.cuda.synchronize(comm.device)
        comm.barrier()   # every rank's slots are zeroed before anyone publishes into them

    def block(self, r0: int, r1: int):
        """This rank's own rows of both maps: two (nz, r1 - r0, nx) views whose [hh] slices are contiguous."""
        return [self.maps[0][:, r0:r1], self.maps[1][:, r0:r1]]

    def peer_ptrs(self, hh: int, r0: int, r1: int, include_self: bool = False):
        """Addresses of rows [r0, r1) of height slice hh inside every other rank's maps (8-byte elements)."""
        if getattr(self, 'mc_base', 0):
            return ([self.mc_base + 8 * ((0 * self.nz + hh) * self.ny + r0) * self.nx], [self.mc_base + 8 * ((1 * self.nz + hh) * self.ny + r0) * self.nx],
                    'multicast')
        wet, hydro = [], []
        for q, b in enumerate(self.base):
            if q == self.comm.rank and not include_self:
                continue
            wet.append(b + 8 * ((0 * self.nz + hh) * self.ny + r0) * self.nx)
            hydro.append(b + 8 * ((1 * self.nz + hh) * self.ny + r0) * self.nx)
        return wet, hydro

    def barrier(self) -> None:
        """Signal-pad barrier on the current CUDA stream: every rank's kernels before it have completed (their peer stores
        included) before any rank's work after it starts."""
        self.hdl.barrier()

    def zero_slice(self, hh: int) -> None:
        """A height slice no rank integrates (no contributing layer at the last output height, delay.py:276-277) is zero in the
        reference (np.zeros, delay.py:248): nobody's kernel writes it, so every rank clears all rows of it in its own maps."""
        self.maps[:, hh].zero_()


def build_cube_ray_sharded(xpts, ypts, zpts, los, model_crs, pts_crs, interpolators, comm: Comm, MAX_SEGMENT_LENGTH=1000.0,
                           MAX_TROPO_HEIGHT=None, gather=True, build_fn=None, host_block=False):
    """``_build_cube_ray`` with the raster row-sharded over ``comm``'s ranks.

    Every rank passes the FULL ``xpts``/``ypts``; rank r integrates rows ``shard_rows(ny, r, world)`` with the global
    reductions hooked in, and (``gather=True``) all ranks return the full ``[wet, hydro]`` (nz, ny, nx) maps as host arrays.
    ``gather='device'`` (nccl) leaves the reassembled maps in HBM and returns them as CUDA tensors -- with ``host_block=True``
    as ``(device_maps, host_rows)``, the rank's own rows having been written to page-locked host memory by the same kernel.
    ``build_fn(xpts, ypts_block, ..., reduce_max=, reduce_sum=)`` defaults to the device path.
    """
    from . import delay as _delay
    from .constants import _ZREF
    ypts = np.asarray(ypts, dtype=np.float64)
    zpts = np.atleast_1d(np.asarray(zpts, dtype=np.float64))
    r0, r1 = shard_rows(ypts.size, comm.rank, comm.world)
    if r1 - r0 < 1:
        raise ValueError(f'raster has {ypts.size} rows: too few to shard over {comm.world} ranks')
    zref = _ZREF if MAX_TROPO_HEIGHT is None else MAX_TROPO_HEIGHT
    if build_fn is None:
        # device path: the row block stays in HBM, the all-gather runs on device tensors (NCCL over NVLink), and the full maps
        # make one trip to page-locked host memory
        import torch
        from ._lib import pinned_empty
        sym = comm.symmetric_maps(zpts.size, ypts.size, np.size(xpts)) if gather else None
        cube = interpolators[0].cube
        prev = _delay._reduce_hooks
        _delay._reduce_hooks = (comm.reduce_max, comm.reduce_sum)
        try:
            if sym is not None:
                # fused step per height (engine._trace_block): K0 -> k_publish (K + 3 words into every rank's exchange slots) ->
                # barrier -> k_plan (MAX / SUM over the slots on the device) -> K3 (stores its rows into every rank's maps) -> barrier.
                # No NCCL collective, no host round trip; everything runs on one side stream the barriers are enqueued on as well
                # (the legacy default stream cannot be handed to the library)
                side = sym.stream
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    cube.h.set_stream(side.cuda_stream)
                    cube.set_exchange(comm.rank, comm.world, sym.xchg_ptrs)
                    try:
                        if host_block:   # this rank's rows also land in page-locked host memory, written by the same kernel
                            local = _delay._build_cube_ray(xpts, ypts[r0:r1], zpts, los, model_crs, pts_crs, interpolators,
                                                           MAX_SEGMENT_LENGTH=MAX_SEGMENT_LENGTH, MAX_TROPO_HEIGHT=zref,
                                                           _peers=lambda hh, a, b: sym.peer_ptrs(hh, r0 + a, r0 + b, include_self=True),
                                                           _exchange=sym, _on_skip=sym.zero_slice)
                        else:
                            local = _delay._build_cube_ray(xpts, ypts[r0:r1], zpts, los, model_crs, pts_crs, interpolators,
                                                           MAX_SEGMENT_LENGTH=MAX_SEGMENT_LENGTH, MAX_TROPO_HEIGHT=zref, _out_arrays=sym.block(r0, r1),
                                                           _peers=lambda hh, a, b: sym.peer_ptrs(hh, r0 + a, r0 + b),
                                                           _exchange=sym, _on_skip=sym.zero_slice)
                    finally:
                        cube.set_exchange(0, 0, None)
                torch.cuda.current_stream().wait_stream(side)
            else:
                local = _delay._build_cube_ray(xpts, ypts[r0:r1], zpts, los, model_crs, pts_crs, interpolators,
                                               MAX_SEGMENT_LENGTH=MAX_SEGMENT_LENGTH, MAX_TROPO_HEIGHT=zref,
                                               _out_device=torch.device('cuda', cube.device))
        finally:
            _delay._reduce_hooks = prev
        if not gather:
            return [a.cpu().numpy() for a in local], (r0, r1)
        if sym is not None and gather == 'device':
            # full maps stay in HBM (valid until the next sharded call of the same shape); with host_block the rank's own rows
            # are returned as host arrays as well
            return ([sym.maps[0], sym.maps[1]], local) if host_block else [sym.maps[0], sym.maps[1]]
        if sym is not None:
            out = [pinned_empty((zpts.size, ypts.size, np.size(xpts))) for _ in range(2)]
            for f in range(2):
                torch.from_numpy(out[f]).copy_(sym.maps[f], non_blocking=True)
            torch.cuda.synchronize()
            return out
        if gather == 'device':   # no symmetric memory: NCCL all-gather, maps stay in HBM
            full = [torch.stack([comm.all_gather_rows(arr[hh], ypts.size) for hh in range(zpts.size)]) for arr in local]
            return (full, [a.cpu().numpy() for a in local]) if host_block else full
        out = [pinned_empty((zpts.size, ypts.size, np.size(xpts))) for _ in local]
        for arr, dst in zip(local, out):
            for hh in range(zpts.size):
                torch.from_numpy(dst[hh]).copy_(comm.all_gather_rows(arr[hh], ypts.size), non_blocking=True)
        torch.cuda.synchronize()
        return out
    else:
        local = build_fn(xpts, ypts[r0:r1], zpts, los, model_crs, pts_crs, interpolators, MAX_SEGMENT_LENGTH=MAX_SEGMENT_LENGTH,
                         MAX_TROPO_HEIGHT=zref, reduce_max=comm.reduce_max, reduce_sum=comm.reduce_sum)
    if not gather:
        return local, (r0, r1)
    out = []
    for arr in local:
        out.append(np.stack([comm.all_gather_rows(arr[hh], ypts.size) for hh in range(zpts.size)], axis=0))
    return out
