import os, sys, torch
sys.path.insert(0, os.getcwd())
from smplfitter_b200 import modeldata
import smplfitter_b200.pt as pt
from torch.profiler import profile, ProfilerActivity
modeldata.use_synthetic_models(True)
bm = pt.BodyModel('smpl', num_betas=10).cuda()
fitter = pt.BodyFitter(bm).cuda()
B = 4096
pose, betas, trans = (torch.randn(B, 72) * 0.2).cuda(), (torch.randn(B, 10) * 0.5).cuda(), torch.randn(B, 3).cuda()
def fb():
    p = pose.detach().requires_grad_(True); b = betas.detach().requires_grad_(True)
    bm(p, b, trans)['vertices'].pow(2).sum().backward()
fb(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    fb(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60))
with torch.no_grad():
    res = bm(pose[:256], betas[:256], trans[:256])
tv, tj = res['vertices'], res['joints']
keys = ['pose_rotvecs', 'shape_betas', 'trans']
def fitb():
    a = tv.detach().requires_grad_(True)
    r = fitter.fit(a, tj, num_iter=3, requested_keys=keys)
    sum(r[k].pow(2).sum() for k in keys).backward()
fitb(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    fitb(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=16, max_name_column_width=60))
print(prof.key_averages().table(sort_by='cpu_time_total', row_limit=10, max_name_column_width=60))
