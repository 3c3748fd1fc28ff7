"""CPU, world_size = 2, gloo: the batch-sharding host logic (scatter -> local fit -> gather)."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smplfitter_b200 import dist as sdist


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fake_fit(tv, tj, scale=1.0):
    # stands in for BodyFitter.fit on CPU: any per-instance function of the inputs
    return {'trans': tv.mean(dim=1) * scale, 'pose_rotvecs': tj.reshape(tj.shape[0], -1) + tv[:, :1, 0]}


def _worker(rank, world, port, total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    V, J = 11, 5
    g = torch.Generator().manual_seed(0)
    tv = torch.randn(total, V, 3, generator=g)
    tj = torch.randn(total, J, 3, generator=g)
    out = sdist.scatter_fit_gather(_fake_fit, total, tv if rank == 0 else None, tj if rank == 0 else None,
                                   V, J, scale=2.0)
    if rank == 0:
        want = _fake_fit(tv, tj, scale=2.0)
        ok = all(torch.allclose(out[k], want[k]) for k in want) and set(out) == set(want)
        q.put(bool(ok))
    else:
        assert out is None
    dist.destroy_process_group()


def test_shard_bounds_cover():
    for total in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [sdist.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_scatter_fit_gather_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_share_beta_hook_symbol_and_view():
    """share_beta over ranks: the C library exports the hook setter, and the zero-copy view class describes the
    buffer the way torch's CUDA array interface expects (the all-reduce itself runs in tests/test_gpu_dist.py)."""
    from smplfitter_b200 import _native

    L = _native.lib()
    assert L.smplfit_set_share_beta_allreduce(_native.ALLREDUCE_FN(0), None, 0) == 0
    cb = _native.ALLREDUCE_FN(lambda p, n, s, u: None)
    assert L.smplfit_set_share_beta_allreduce(cb, None, 0) != 0  # a hook needs the global batch size
    assert L.smplfit_set_share_beta_allreduce(cb, None, 8) == 0
    assert L.smplfit_set_share_beta_allreduce(_native.ALLREDUCE_FN(0), None, 0) == 0
    v = sdist._RawDoubles(0x1000, 65)
    assert v.__cuda_array_interface__['shape'] == (65,) and v.__cuda_array_interface__['typestr'] == '<f8'
