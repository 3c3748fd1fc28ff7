import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smplfitter_b200.pt import BodyFitter, BodyModel  # noqa: E402
B = 4096
bm = BodyModel('smpl').cuda(); fitter = BodyFitter(bm).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
fw = bm(torch.randn(B, 72, device='cuda', generator=g) * 0.1, torch.randn(B, 10, device='cuda', generator=g) * 0.5, torch.randn(B, 3, device='cuda', generator=g))
h_tv = torch.empty(fw['vertices'].shape, pin_memory=True).copy_(fw['vertices'].cpu()); h_tj = torch.empty(fw['joints'].shape, pin_memory=True).copy_(fw['joints'].cpu())
kw = dict(num_iter=3, beta_regularizer=1.0, requested_keys=['pose_rotvecs', 'shape_betas'])
cs = int(sys.argv[1])
for _ in range(3):
    fitter.fit_from_host(h_tv, h_tj, chunk_size=cs, **kw)
