"""Batch sharding of the fit over the GPUs of one node.

Instances are independent, so the only communication is moving inputs out of and results back
into one rank: no collective inside the iteration.  Works with any ``torch.distributed`` backend
(NCCL over NVLink on GPUs; gloo in the CPU tests, where the local fit is injected).

``scatter_fit_gather`` pipelines the three phases: the source rank posts the point-to-point sends chunk by
chunk (NCCL has no native scatterv), every rank starts fitting chunk *c* as soon as it has arrived while chunk
*c + 1* is still on the wire, and the per-instance results (a few hundred bytes each) travel back as ONE packed
row block per rank.
"""

from __future__ import annotations

import contextlib
import os
from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, near-equal shard [lo, hi) of ``total`` items for ``rank``."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _peer(group, group_rank: int) -> int:
    """Point-to-point calls take GLOBAL ranks; shard indices are group-local."""
    return group_rank if group is None else dist.get_global_rank(group, group_rank)


def bind_rank_to_local_cpus(local_rank: int) -> Optional[list]:
    """Pin the calling process to the CPU cores NVML reports as local to GPU ``local_rank`` (their NUMA node), so
    that pinned staging buffers allocated afterwards are NUMA-local to the GPU that reads them.  When every GPU
    reports the same core set (a single-socket or virtualised host) the set is split evenly between the local
    ranks instead.  Returns the chosen cores (None when NVML or affinity control is unavailable)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        n = pynvml.nvmlDeviceGetCount()
        words = (os.cpu_count() + 63) // 64

        def cores(i):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            return [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]

        mine = cores(local_rank % n)
        allowed = sorted(set(mine) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        sharing = [i for i in range(n) if cores(i) == mine]
        if len(sharing) > 1 and len(allowed) >= 2 * len(sharing):
            k = sharing.index(local_rank % n)
            per = len(allowed) // len(sharing)
            allowed = allowed[k * per:(k + 1) * per]
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def scatter_rows(x: Optional[torch.Tensor], total: int, src: int = 0, group=None, device=None,
                 tail: tuple = (), dtype=torch.float32) -> torch.Tensor:
    """Scatter the rows of ``x`` (held by group rank ``src``; other ranks pass None) by ``shard_bounds``."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(total, world, rank)
    if rank == src:
        reqs = []
        for r in range(world):
            a, b = shard_bounds(total, world, r)
            if r != src and b > a:  # empty shards (total < world) are not sent
                reqs.append(dist.isend(x[a:b].contiguous(), _peer(group, r), group=group))
        out = x[lo:hi].to(device=device, dtype=dtype) if device is not None else x[lo:hi].to(dtype=dtype)
        for q in reqs:
            q.wait()
        return out
    out = torch.empty((hi - lo, *tail), dtype=dtype, device=device)
    if hi > lo:
        dist.recv(out, _peer(group, src), group=group)
    return out


def gather_rows(x: torch.Tensor, total: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Inverse of ``scatter_rows``: concatenate the shards on group rank ``dst`` (None elsewhere)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rank == dst:
        out = torch.empty((total, *x.shape[1:]), dtype=x.dtype, device=x.device)
        reqs = []
        for r in range(world):
            lo, hi = shard_bounds(total, world, r)
            if r == dst:
                out[lo:hi].copy_(x)
            elif hi > lo:
                reqs.append(dist.irecv(out[lo:hi], _peer(group, r), group=group))
        for q in reqs:
            q.wait()
        return out
    if x.shape[0] > 0:
        dist.send(x.contiguous(), _peer(group, dst), group=group)
    return None


class _RawDoubles:
    """Zero-copy view of `count` device doubles at `ptr` for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {'data': (int(ptr), False), 'shape': (int(count),), 'typestr': '<f8',
                                         'version': 2, 'strides': None}


class share_beta_across_ranks:
    """Context manager: while active, every ``share_beta=True`` shape solve of this process sums its centred normal
    equations over the ranks of ``group`` (ONE all-reduce of S(S+1)/2 + S doubles per solve -- the only cross-instance
    term of the fit, pt/lstsq.py:24-26), so the batch shards of all ranks share one set of betas.  Every rank must run
    the same number of solves (same ``num_iter``, one ``fit`` call each).  ``global_batch`` = instances over all ranks."""

    def __init__(self, global_batch: int, group=None, device=None):
        from . import _native

        self._native = _native
        self.total, self.group, self.device = int(global_batch), group, device
        self.error = None

        def hook(ptr, count, stream, user):
            try:
                t = torch.as_tensor(_RawDoubles(ptr, count), device=self.device)
                dist.all_reduce(t, group=self.group)
            except BaseException as e:  # an exception cannot cross the C frame: keep it for __exit__
                self.error = e

        self._cb = _native.ALLREDUCE_FN(hook)  # keep the callback object alive while installed

    def __enter__(self):
        self._native.check(self._native.lib().smplfit_set_share_beta_allreduce(self._cb, None, self.total))
        return self

    def __exit__(self, *exc):
        self._native.lib().smplfit_set_share_beta_allreduce(self._native.ALLREDUCE_FN(0), None, 0)
        if self.error is not None and exc[0] is None:
            raise self.error
        return False


def _pack(res: dict, keys: list, n: int) -> torch.Tensor:
    return torch.cat([res[k].reshape(n, -1) for k in keys], dim=1) if keys else torch.empty((n, 0))


def scatter_fit_gather(fit_fn: Callable[..., dict], total: int, target_vertices: Optional[torch.Tensor],
                       target_joints: Optional[torch.Tensor], num_vertices: int, num_joints: int,
                       has_joints: bool = True, src: int = 0, group=None, device=None, n_chunks: int = 2,
                       **fit_kwargs) -> Optional[dict]:
    """Shard a batch held by group rank ``src`` over the group, run ``fit_fn(verts, joints, **kw)`` on
    each shard (``BodyFitter.fit`` of the rank-local fitter) and gather the result dictionary
    on ``src``.  Every shard travels in ``n_chunks`` pieces; a rank fits piece *c* while piece *c + 1* is still
    arriving.  ``share_beta=True`` couples the instances of the whole batch: the shards are then fitted in one
    piece each under ``share_beta_across_ranks`` (one all-reduce of the normal equations per shape solve); it needs
    a CUDA fitter (the hook lives in the C library) and at least one instance per rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(total, world, rank)
    n = hi - lo
    shared = bool(fit_kwargs.get('share_beta'))
    if shared:
        if total < world:
            raise ValueError('share_beta over ranks needs at least one instance per rank')
        n_chunks = 1
    n_chunks = max(1, min(int(n_chunks), max(n, 1)))
    V, J = num_vertices, num_joints

    def pieces(r):
        a, b = shard_bounds(total, world, r)
        return [(a + s, a + e) for s, e in (shard_bounds(b - a, n_chunks, c) for c in range(n_chunks)) if e > s]

    mine = pieces(rank)
    waits = []  # per local piece: requests that must complete before it can be fitted
    if rank == src:
        reqs = []
        for c in range(n_chunks):  # chunk-major, so every receiver gets its first piece early
            for r in range(world):
                if r == src:
                    continue
                pr = pieces(r)
                if c < len(pr):
                    a, b = pr[c]
                    reqs.append(dist.isend(target_vertices[a:b], _peer(group, r), group=group))
                    if has_joints:
                        reqs.append(dist.isend(target_joints[a:b], _peer(group, r), group=group))
        tv_loc = target_vertices[lo:hi]
        tj_loc = target_joints[lo:hi] if has_joints else None
        waits = [[] for _ in mine]
    else:
        reqs = []
        tv_loc = torch.empty((n, V, 3), dtype=torch.float32, device=device)
        tj_loc = torch.empty((n, J, 3), dtype=torch.float32, device=device) if has_joints else None
        for a, b in mine:
            w = [dist.irecv(tv_loc[a - lo:b - lo], _peer(group, src), group=group)]
            if has_joints:
                w.append(dist.irecv(tj_loc[a - lo:b - lo], _peer(group, src), group=group))
            waits.append(w)
    parts = []
    ctx = share_beta_across_ranks(total, group, device) if shared else contextlib.nullcontext()
    with ctx:
        for (a, b), w in zip(mine, waits):
            for q in w:
                q.wait()  # NCCL: orders the current stream after the transfer, does not block the host
            parts.append(fit_fn(tv_loc[a - lo:b - lo], tj_loc[a - lo:b - lo] if has_joints else None, **fit_kwargs))
    if not parts:  # empty shard: the result keys / shapes still come from the fit function
        parts.append(fit_fn(tv_loc, tj_loc, **fit_kwargs))
    keys = sorted(parts[0])
    local = {k: (torch.cat([p[k] for p in parts]) if len(parts) > 1 else parts[0][k]) for k in keys}
    packed = _pack(local, keys, n)
    g = gather_rows(packed.contiguous(), total, src, group)
    for q in reqs:
        q.wait()
    if rank != src:
        return None
    out, off = {}, 0
    for k in keys:
        tail = tuple(local[k].shape[1:])
        width = 1
        for s in tail:
            width *= s
        out[k] = g[:, off:off + width].reshape(total, *tail)
        off += width
    return out
