#!/usr/bin/env python
"""Is the end-to-end time bound by the copy or by the fits?  fit_from_host with lighter / heavier fits."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smplfitter_b200.pt import BodyFitter, BodyModel  # noqa: E402
B = 4096
bm = BodyModel('smpl').cuda(); fitter = BodyFitter(bm).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
fw = bm(torch.randn(B, 72, device='cuda', generator=g) * 0.1, torch.randn(B, 10, device='cuda', generator=g) * 0.5, torch.randn(B, 3, device='cuda', generator=g))
tv, tj = fw['vertices'], fw['joints']
h_tv = torch.empty(tv.shape, pin_memory=True).copy_(tv.cpu()); h_tj = torch.empty(tj.shape, pin_memory=True).copy_(tj.cpu())
def timed(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
out = {}
for it, adj in ((1, False), (1, True), (2, True), (3, True)):
    kw = dict(num_iter=it, final_adjust_rots=adj, beta_regularizer=1.0, requested_keys=['pose_rotvecs', 'shape_betas'])
    out[f'resident_it{it}_adj{int(adj)}'] = timed(lambda: fitter.fit(tv, tj, **kw))
    for cs in (512, 1024):
        out[f'host_it{it}_adj{int(adj)}_c{cs}'] = timed(lambda: fitter.fit_from_host(h_tv, h_tj, chunk_size=cs, **kw))
print(json.dumps({k: round(v, 3) for k, v in out.items()}, indent=1))
