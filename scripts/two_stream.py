#!/usr/bin/env python
"""Experiment: one fit of 4096 vs the same batch as K sub-batches on K streams (do the short latency-bound stages of
one sub-batch hide under the vertex passes of another?)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smplfitter_b200.pt import BodyFitter, BodyModel  # noqa: E402

B = 4096
bm = BodyModel('smpl').cuda()
fitter = BodyFitter(bm).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
pose = torch.randn(B, 72, device='cuda', generator=g) * 0.1
betas = torch.randn(B, 10, device='cuda', generator=g) * 0.5
trans = torch.randn(B, 3, device='cuda', generator=g)
fw = bm(pose, betas, trans)
tv, tj = fw['vertices'], fw['joints']
kw = dict(num_iter=3, beta_regularizer=1.0, requested_keys=['pose_rotvecs', 'shape_betas'])


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


print('single 4096: %.3f ms' % timed(lambda: fitter.fit(tv, tj, **kw)))
for K in (2, 3, 4):
    streams = [torch.cuda.Stream() for _ in range(K)]
    n = B // K
    parts = [(tv[i * n:(i + 1) * n].contiguous(), tj[i * n:(i + 1) * n].contiguous()) for i in range(K)]

    def run():
        cur = torch.cuda.current_stream()
        for s, (a, b) in zip(streams, parts):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                fitter.fit(a, b, **kw)
        for s in streams:
            cur.wait_stream(s)

    print('%d x %d on %d streams: %.3f ms' % (K, n, K, timed(run)))
    def run_serial():
        for (a, b) in parts:
            fitter.fit(a, b, **kw)
    print('%d x %d serial: %.3f ms' % (K, n, timed(run_serial)))
