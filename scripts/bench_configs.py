#!/usr/bin/env python
"""Timing of the other BASELINE.json configs (not the headline line of bench.py):
SMPL-X fit (config 3), 1024-vertex-subset fit at batch 16384 (config 4), SMPL->SMPL-X conversion
(config 5), forward LBS.  Prints one JSON object; run under gpurun on one GPU."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smplfitter_b200 import _native  # noqa: E402
from smplfitter_b200.pt import BodyConverter, BodyFitter, BodyModel  # noqa: E402


def timed(fn, steps=5, warm=2):
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


def fit_case(model, B, mkw=None, **fkw):
    bm = BodyModel(model, **(mkw or {})).cuda()
    fitter = BodyFitter(bm).cuda()
    g = torch.Generator(device='cuda').manual_seed(1)
    pose = torch.randn(B, 3 * bm.num_joints, device='cuda', generator=g) * 0.1
    betas = torch.randn(B, bm.num_betas, device='cuda', generator=g) * 0.5
    trans = torch.randn(B, 3, device='cuda', generator=g)
    fw = bm(pose, betas, trans)
    kw = dict(num_iter=3, beta_regularizer=1.0, requested_keys=['pose_rotvecs', 'shape_betas'])
    kw.update(fkw)
    ms = timed(lambda: fitter.fit(fw['vertices'], fw['joints'], **kw))
    _native.profile(True)
    fitter.fit(fw['vertices'], fw['joints'], **kw)
    torch.cuda.synchronize()
    _native.profile(False)
    prof = {k: round(v[1], 3) for k, v in sorted(_native.profile_report().items(), key=lambda kv: -kv[1][1])[:6]}
    fit = fitter.fit(fw['vertices'], fw['joints'], **kw)
    re = bm(fit['pose_rotvecs'], fit['shape_betas'], fit['trans'])
    v2v = (re['vertices'] - fw['vertices']).norm(dim=-1).mean().item() * 1000
    fwd_ms = timed(lambda: bm(pose, betas, trans))
    _native.profile(True)
    bm(pose, betas, trans)
    torch.cuda.synchronize()
    _native.profile(False)
    fprof = {k: round(v[1], 3) for k, v in _native.profile_report().items()}
    return {'model': model, 'B': B, 'V': bm.num_vertices, 'J': bm.num_joints, 'S': bm.num_betas, 'ms': ms,
            'fits_per_s': B / ms * 1000, 'roundtrip_v2v_mm': v2v, 'top_kernels_ms': prof,
            'forward_ms': fwd_ms, 'forwards_per_s': B / fwd_ms * 1000, 'forward_kernels_ms': fprof}


def converter_case(B):
    import scipy.sparse as sp

    bm_in, bm_out = BodyModel('smpl').cuda(), BodyModel('smplx').cuda()
    rs = np.random.RandomState(0)
    vin, vout = bm_in.num_vertices, bm_out.num_vertices
    # synthetic barycentric transfer: each SMPL-X vertex from 3 nearby SMPL template vertices
    tin = bm_in._t_template_mesh.cpu().numpy()
    tout = bm_out._t_template_mesh.cpu().numpy()
    cols = np.empty((vout, 3), np.int64)
    for lo in range(0, vout, 512):
        d = ((tout[lo:lo + 512, None] - tin[None]) ** 2).sum(-1)
        cols[lo:lo + 512] = np.argsort(d, axis=1)[:, :3]
    w = rs.dirichlet([2, 2, 2], size=vout).astype(np.float32)
    m = sp.csr_matrix((w.reshape(-1), (np.repeat(np.arange(vout), 3), cols.reshape(-1))), shape=(vout, vin))
    conv = BodyConverter(bm_in, bm_out, vertex_converter_csr=m).cuda()
    g = torch.Generator(device='cuda').manual_seed(2)
    pose = torch.randn(B, 72, device='cuda', generator=g) * 0.1
    betas = torch.randn(B, 10, device='cuda', generator=g) * 0.5
    trans = torch.randn(B, 3, device='cuda', generator=g)
    ms = timed(lambda: conv.convert(pose, betas, trans, num_iter=1), steps=3, warm=1)
    return {'B': B, 'ms': ms, 'conversions_per_s': B / ms * 1000}


out = {}
which = sys.argv[1:] or ['smplx', 'subset', 'converter']
if 'smpl' in which:
    out['smpl_4096'] = fit_case('smpl', 4096)
if 'smplx' in which:
    out['smplx_4096'] = fit_case('smplx', 4096)
if 'subset' in which:
    out['subset1024_16384'] = fit_case('smpl', 16384, dict(vertex_subset_size=1024))
if 'converter' in which:
    out['converter_smpl_to_smplx_4096'] = converter_case(4096)
print(json.dumps(out))
