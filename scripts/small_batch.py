#!/usr/bin/env python
"""One fit at a small batch (for an ncu launch list): python scripts/small_batch.py B [model]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smplfitter_b200.pt import BodyFitter, BodyModel  # noqa: E402

B = int(sys.argv[1])
model = sys.argv[2] if len(sys.argv) > 2 else 'smpl'
bm = BodyModel(model).cuda()
fitter = BodyFitter(bm).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
pose = torch.randn(B, 3 * bm.num_joints, device='cuda', generator=g) * 0.1
betas = torch.randn(B, bm.num_betas, device='cuda', generator=g) * 0.5
trans = torch.randn(B, 3, device='cuda', generator=g)
fw = bm(pose, betas, trans)
kw = dict(num_iter=3, beta_regularizer=1.0, requested_keys=['pose_rotvecs', 'shape_betas'])
for _ in range(3):
    fitter.fit(fw['vertices'], fw['joints'], **kw)
torch.cuda.synchronize()
