"""Worker of tests/test_gpu_variants.py: runs forward + fit on two tiny models with whatever kernel-variant
environment the parent set, and saves the results."""
import sys

import numpy as np
import torch

from smplfitter_b200 import modeldata
from smplfitter_b200.pt import BodyFitter, BodyModel

modeldata.use_synthetic_models(True)

out = {}
for mname, B in (('smpl_tiny', 45), ('smplx_tiny', 33)):
    bm = BodyModel(mname).cuda()
    fitter = BodyFitter(bm).cuda()
    rs = np.random.RandomState(5)
    pose = torch.from_numpy((rs.randn(B, 3 * bm.num_joints) * 0.2).astype(np.float32)).cuda()
    betas = torch.from_numpy((rs.randn(B, bm.num_betas) * 0.5).astype(np.float32)).cuda()
    trans = torch.from_numpy(rs.randn(B, 3).astype(np.float32)).cuda()
    fw = bm(pose, betas, trans)
    out[mname + '_vertices'] = fw['vertices'].cpu().numpy()
    tv = fw['vertices'] + torch.from_numpy((rs.randn(B, bm.num_vertices, 3) * 0.003).astype(np.float32)).cuda()
    for tag, kw in (('joints', dict(target_joints=fw['joints'])), ('nojoints', {}),
                    ('weights', dict(target_joints=fw['joints'],
                                     vertex_weights=torch.from_numpy(rs.rand(B, bm.num_vertices).astype(np.float32)).cuda(),
                                     joint_weights=torch.from_numpy(rs.rand(B, bm.num_joints).astype(np.float32)).cuda()))):
        fit = fitter.fit(tv, num_iter=2, beta_regularizer=1.0, requested_keys=['pose_rotvecs', 'shape_betas'], **kw)
        for k in ('pose_rotvecs', 'shape_betas', 'trans'):
            out[f'{mname}_{tag}_{k}'] = fit[k].cpu().numpy()
np.savez(sys.argv[1], **out)
