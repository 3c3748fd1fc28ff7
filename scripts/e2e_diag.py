#!/usr/bin/env python
"""Where the end-to-end (host-resident inputs) time goes: pinned H2D bandwidth of this box, then
fit_from_host at several chunk sizes against the resident fit.  Run under gpurun on one GPU."""
import json
import os
import sys
import time

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
h_tv = torch.empty(tv.shape, pin_memory=True).copy_(tv.cpu())
h_tj = torch.empty(tj.shape, pin_memory=True).copy_(tj.cpu())
kw = dict(num_iter=3, beta_regularizer=1.0, requested_keys=['pose_rotvecs', 'shape_betas'])


def timed(fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    t_issue = (time.perf_counter() - t0) / steps * 1000
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, t_issue


out = {'pinned': bool(h_tv.is_pinned())}
d = torch.empty_like(tv)
ms, _ = timed(lambda: d.copy_(h_tv, non_blocking=True))
out['h2d_gbs_340MB'] = h_tv.numel() * 4 / ms / 1e6
ms, _ = timed(lambda: [d[i:i + 512].copy_(h_tv[i:i + 512], non_blocking=True) for i in range(0, B, 512)])
out['h2d_gbs_chunks512'] = h_tv.numel() * 4 / ms / 1e6
hb = torch.empty(tv.shape, pin_memory=True)
ms, _ = timed(lambda: hb.copy_(d, non_blocking=True))
out['d2h_gbs'] = h_tv.numel() * 4 / ms / 1e6
ms, iss = timed(lambda: fitter.fit(tv, tj, **kw))
out['resident_ms'], out['resident_issue_ms'] = ms, iss
for cs in (128, 256, 512, 1024, 2048, 4096):
    ms, iss = timed(lambda: fitter.fit_from_host(h_tv, h_tj, chunk_size=cs, **kw))
    out[f'from_host_{cs}_ms'], out[f'from_host_{cs}_issue_ms'] = ms, iss
for bs in (32, 256, 1024):
    ms, iss = timed(lambda: fitter.fit(tv[:bs], tj[:bs], **kw), steps=10)
    out[f'resident_B{bs}_ms'], out[f'resident_B{bs}_issue_ms'] = ms, iss
print(json.dumps(out, indent=1))
