"""Batch sharding of the fit over the GPUs of one node.

Instances are independent, so the only communication is moving inputs out of and results back
into one rank: no collective inside the iteration.  Works with any ``torch.distributed`` backend
(NCCL on GPUs; gloo in the CPU tests, where the local fit is injected).
"""

from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, near-equal shard [lo, hi) of ``total`` items for ``rank``."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scatter_rows(x: Optional[torch.Tensor], total: int, src: int = 0, group=None, device=None,
                 tail: tuple = (), dtype=torch.float32) -> torch.Tensor:
    """Scatter the rows of ``x`` (held by ``src``; other ranks pass None) by ``shard_bounds``."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(total, world, rank)
    out = torch.empty((hi - lo, *tail), dtype=dtype, device=device)
    if rank == src:
        chunks = [x[slice(*shard_bounds(total, world, r))].contiguous() for r in range(world)]
        # unequal shards: point-to-point sends (NCCL has no native scatterv)
        reqs = [dist.isend(chunks[r], r, group=group) for r in range(world) if r != src]
        out.copy_(chunks[src])
        for q in reqs:
            q.wait()
    else:
        dist.recv(out, src, group=group)
    return out


def gather_rows(x: torch.Tensor, total: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Inverse of ``scatter_rows``: concatenate the shards on ``dst`` (None elsewhere)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rank == dst:
        out = torch.empty((total, *x.shape[1:]), dtype=x.dtype, device=x.device)
        reqs = []
        for r in range(world):
            lo, hi = shard_bounds(total, world, r)
            if r == dst:
                out[lo:hi].copy_(x)
            else:
                reqs.append(dist.irecv(out[lo:hi], r, group=group))
        for q in reqs:
            q.wait()
        return out
    dist.send(x.contiguous(), dst, group=group)
    return None


def scatter_fit_gather(fit_fn: Callable[..., dict], total: int, target_vertices: Optional[torch.Tensor],
                       target_joints: Optional[torch.Tensor], num_vertices: int, num_joints: int,
                       has_joints: bool = True, src: int = 0, group=None, device=None, **fit_kwargs) -> Optional[dict]:
    """Shard a batch held by rank ``src`` over the group, run ``fit_fn(verts, joints, **kw)`` on
    each shard (``BodyFitter.fit`` of the rank-local fitter) and gather the result dictionary
    on ``src``.  ``share_beta`` couples the instances of the whole batch (one all-reduce of the normal
    equations per shape solve would be needed) and is rejected here."""
    if fit_kwargs.get('share_beta'):
        raise NotImplementedError('share_beta couples all instances: fit it on one rank (BodyFitter.fit)')
    tv = scatter_rows(target_vertices, total, src, group, device, (num_vertices, 3))
    tj = scatter_rows(target_joints, total, src, group, device, (num_joints, 3)) if has_joints else None
    local = fit_fn(tv, tj, **fit_kwargs)
    out = {}
    for k in sorted(local):
        g = gather_rows(local[k].contiguous(), total, src, group)
        if g is not None:
            out[k] = g
    return out if dist.get_rank(group) == src else None
