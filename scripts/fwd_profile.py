#!/usr/bin/env python
"""Per-kernel device times of BodyModel.forward (CUDA events around every launch) at the BASELINE batch size.
    python scripts/fwd_profile.py [model] [batch] [--short]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smplfitter_b200 import _native, modeldata  # noqa: E402
from smplfitter_b200.pt import BodyModel  # noqa: E402

modeldata.use_synthetic_models(True)
args = [a for a in sys.argv[1:] if not a.startswith('--')]
model = args[0] if args else 'smpl'
B = int(args[1]) if len(args) > 1 else 4096
short = '--short' in sys.argv
bm = BodyModel(model).cuda()
rs = np.random.RandomState(42)
pose = torch.from_numpy((rs.randn(B, 3 * bm.num_joints) * 0.1).astype(np.float32)).cuda()
betas = torch.from_numpy((rs.randn(B, bm.num_betas) * 0.5).astype(np.float32)).cuda()
trans = torch.from_numpy(rs.randn(B, 3).astype(np.float32)).cuda()
for _ in range(3):
    bm(pose, betas, trans)
torch.cuda.synchronize()
if short:
    bm(pose, betas, trans)
    torch.cuda.synchronize()
    sys.exit(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    bm(pose, betas, trans)
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1) / 20
_native.profile(True)
for _ in range(10):
    bm(pose, betas, trans)
torch.cuda.synchronize()
_native.profile(False)
prof = {k: v[1] / 10 for k, v in _native.profile_report().items()}
V = bm.num_vertices
print(json.dumps({'model': model, 'B': B, 'ms_per_call': total, 'kernels_ms': prof,
                  'store_gbs_of_fused': B * V * 12 / (prof.get('k_fwd_fused', 1e9) / 1000) / 1e9}))
