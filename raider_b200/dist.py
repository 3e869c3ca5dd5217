"""Raster sharding over the GPUs of one node: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch).

The reference has no parallel delay path (``nproc > 1`` raises NotImplementedError, tools/RAiDER/delay.py:178-185).
Rays are independent except for three whole-raster quantities, which is why naive tiling changes results
(SURVEY.md section 0 facts 4-5):

* ``nParts[k] = ceil(max_over_raster(ray_length[k]) / MAX_SEGMENT_LENGTH) + 1``      (delay.py:283)
* the ``.all()`` clamp predicates on the sample heights                              (delay.py:306-311)
* the ``isnan(ray_lengths).all()`` convergence check                                 (delay.py:279)

So the path shards as: contiguous row blocks of the query raster per rank (cube replicated, it is MBs), K0 on each
rank, ONE all-reduce(MAX) of K doubles + ONE all-reduce(SUM) of 3 counters, K3 on each rank, then an all-gather of the
two output row blocks so every rank holds the full delay map.  There is no other data-path collective.

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


def build_cube_ray_sharded(xpts, ypts, zpts, los, model_crs, pts_crs, interpolators, comm: Comm, MAX_SEGMENT_LENGTH=1000.0,
                           MAX_TROPO_HEIGHT=None, gather=True, build_fn=None):
    """``_build_cube_ray`` with the raster row-sharded over ``comm``'s ranks.

    Every rank passes the FULL ``xpts``/``ypts``; rank r integrates rows ``shard_rows(ny, r, world)`` with the global
    reductions hooked in, and (``gather=True``) all ranks return the full ``[wet, hydro]`` (nz, ny, nx) maps.
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
        prev = _delay._reduce_hooks
        _delay._reduce_hooks = (comm.reduce_max, comm.reduce_sum)
        try:
            local = _delay._build_cube_ray(xpts, ypts[r0:r1], zpts, los, model_crs, pts_crs, interpolators,
                                           MAX_SEGMENT_LENGTH=MAX_SEGMENT_LENGTH, MAX_TROPO_HEIGHT=zref,
                                           _out_device=torch.device('cuda', interpolators[0].cube.device))
        finally:
            _delay._reduce_hooks = prev
        if not gather:
            return [a.cpu().numpy() for a in local], (r0, r1)
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
