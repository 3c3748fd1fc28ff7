"""Forward LBS timing (B = 4096 SMPL, CUDA events, outputs larger than L2): one JSON object; not a bench line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smplfitter_b200 import modeldata  # noqa: E402
import smplfitter_b200.pt as pt  # noqa: E402

modeldata.use_synthetic_models(True)
out = {}
for name in ('smpl', 'smplx'):
    bm = pt.BodyModel(name).cuda()
    B = 4096
    pose = (torch.randn(B, 3 * bm.num_joints) * 0.2).cuda()
    betas, trans = (torch.randn(B, bm.num_betas) * 0.5).cuda(), torch.randn(B, 3).cuda()
    with torch.no_grad():
        for _ in range(5):
            bm(pose, betas, trans)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            bm(pose, betas, trans)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    out[name] = {'ms_per_call': round(ms, 4), 'store_gbs': round(B * bm.num_vertices * 12 / ms / 1e6, 1)}
out['env'] = os.environ.get('SMPLFIT_B200_FWD_EW', 'default')
print(json.dumps(out))
