"""Cost of the gradient path on the GPU box: fit / forward with and without a backward (CUDA events, after warm-up).
Writes one JSON object to stdout; not a bench line."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smplfitter_b200 import modeldata  # noqa: E402
import smplfitter_b200.pt as pt  # noqa: E402

modeldata.use_synthetic_models(True)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    bm = pt.BodyModel('smpl', num_betas=10).cuda()
    fitter = pt.BodyFitter(bm).cuda()
    out = {}
    for B in (32, 256, 1024):
        torch.manual_seed(0)
        with torch.no_grad():
            res = bm((torch.randn(B, 72) * 0.2).cuda(), (torch.randn(B, 10) * 0.5).cuda(), torch.randn(B, 3).cuda())
        tv, tj = res['vertices'], res['joints']
        keys = ['pose_rotvecs', 'shape_betas', 'trans']

        def fwd():
            with torch.no_grad():
                fitter.fit(tv, tj, num_iter=3, requested_keys=keys)

        def fwd_bwd():
            a = tv.detach().requires_grad_(True)
            r = fitter.fit(a, tj, num_iter=3, requested_keys=keys)
            sum(r[k].pow(2).sum() for k in keys).backward()

        out[f'fit_B{B}'] = {'fit_ms': round(timed(fwd, 5), 3), 'fit_plus_backward_ms': round(timed(fwd_bwd, 2), 1),
                            'peak_mem_GB': round(torch.cuda.max_memory_allocated() / 2**30, 2)}
    B = 4096
    pose, betas, trans = (torch.randn(B, 72) * 0.2).cuda(), (torch.randn(B, 10) * 0.5).cuda(), torch.randn(B, 3).cuda()

    def f():
        with torch.no_grad():
            bm(pose, betas, trans)

    def fb():
        p = pose.detach().requires_grad_(True)
        bm(p, betas, trans)['vertices'].pow(2).sum().backward()

    out['forward_B4096'] = {'forward_ms': round(timed(f, 10), 3), 'forward_plus_backward_ms': round(timed(fb, 3), 1)}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
