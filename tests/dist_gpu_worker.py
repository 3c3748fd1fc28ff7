"""Rank program of tests/test_gpu_dist.py (run under torchrun, one rank per GPU, NCCL): the global batch lives on
rank 0's GPU, is scattered chunk by chunk, fitted by every rank's BodyFitter, and the results are gathered on
rank 0, which checks them against its own single-GPU fit of the whole batch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from smplfitter_b200 import dist as sdist, modeldata  # noqa: E402
from smplfitter_b200.pt import BodyFitter, BodyModel  # noqa: E402


def main():
    modeldata.use_synthetic_models(True)
    rank, local, world = (int(os.environ[k]) for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 70
    bm = BodyModel('smpl_tiny').to(dev)
    fitter = BodyFitter(bm).to(dev)
    J, S, V = bm.num_joints, bm.num_betas, bm.num_vertices
    kw = dict(num_iter=2, beta_regularizer=1.0, requested_keys=['pose_rotvecs', 'shape_betas'])
    tv = tj = None
    if rank == 0:
        rs = np.random.RandomState(5)
        fw = bm(torch.from_numpy((rs.randn(total, 3 * J) * 0.2).astype(np.float32)).to(dev),
                torch.from_numpy((rs.randn(total, S) * 0.5).astype(np.float32)).to(dev),
                torch.from_numpy(rs.randn(total, 3).astype(np.float32)).to(dev))
        tv, tj = fw['vertices'].contiguous(), fw['joints'].contiguous()
    out = sdist.scatter_fit_gather(fitter.fit, total, tv, tj, V, J, device=dev, n_chunks=2, **kw)
    res = {'rank': rank, 'ok': True}
    if rank == 0:
        want = fitter.fit(tv, tj, **kw)
        res['keys'] = sorted(out) == sorted(want)
        res['diff'] = {k: float((out[k] - want[k]).abs().max().item()) for k in want}
        # instances are independent and every rank runs the same kernels: the sharded result is the single-GPU one
        res['ok'] = res['keys'] and all(v <= 1e-6 for v in res['diff'].values())
    else:
        assert out is None
    # share_beta over the ranks: one all-reduce of the centred normal equations per shape solve
    if total >= world:
        kws = dict(kw, share_beta=True)
        outs = sdist.scatter_fit_gather(fitter.fit, total, tv, tj, V, J, device=dev, **kws)
        if rank == 0:
            wants = fitter.fit(tv, tj, **kws)
            res['diff_shared'] = {k: float((outs[k] - wants[k]).abs().max().item()) for k in wants}
            res['betas_equal'] = bool((outs['shape_betas'] - outs['shape_betas'][:1]).abs().max().item() == 0.0)
            res['ok'] = res['ok'] and res['betas_equal'] and all(v <= 2e-5 for v in res['diff_shared'].values())
    if rank == 0:
        print('DIST_RESULT ' + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
