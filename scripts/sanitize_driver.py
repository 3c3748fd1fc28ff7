"""Small workload for compute-sanitizer (scripts/gpu_sanitize.sh): one call of every kernel family on tiny models, with
ragged batch sizes (not multiples of 32 / 128) so that the tail handling of the TMA rings and tcgen05 tiles runs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smplfitter_b200 import modeldata  # noqa: E402
from smplfitter_b200.pt import BodyConverter, BodyFitter, BodyModel  # noqa: E402

modeldata.use_synthetic_models(True)
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
dev = torch.device('cuda', 0)
rs = np.random.RandomState(0)


def params(B, J, S):
    f = lambda *s: torch.from_numpy(rs.randn(*s).astype(np.float32)).to(dev)  # noqa: E731
    return f(B, 3 * J) * 0.2, f(B, S) * 0.5, f(B, 3)


def run(name, fn):
    if which in ('all', name):
        fn()
        torch.cuda.synchronize()
        print('ok', name, flush=True)


bm = BodyModel('smpl_tiny').to(dev)
J, S, V = bm.num_joints, bm.num_betas, bm.num_vertices
fitter = BodyFitter(bm).to(dev)


def fit_ragged():
    for B in (6, 70, 77, 131):
        p, b, t = params(B, J, S)
        fw = bm(p, b, t)
        fitter.fit(fw['vertices'], fw['joints'], num_iter=2, beta_regularizer=1.0, requested_keys=['pose_rotvecs'])


def fit_options():
    B = 37
    p, b, t = params(B, J, S)
    fw = bm(p, b, t)
    vw = torch.rand(B, V, device=dev) + 0.2
    jw = torch.rand(B, J, device=dev) + 0.2
    fitter.fit(fw['vertices'], None, num_iter=2)                                       # regressed joints (aT_out path)
    fitter.fit(fw['vertices'], fw['joints'], vw, jw, num_iter=2)                       # weighted (general shape pass)
    fitter.fit(fw['vertices'], fw['joints'], num_iter=2, scale_target=True)            # scale pass of the final solve
    fitter.fit(fw['vertices'], fw['joints'], num_iter=2, share_beta=True)
    fitter.fit(fw['vertices'], fw['joints'], num_iter=2, share_beta=True, scale_fit=True)  # partial share
    for _ in range(3):  # repeated identical call: captured into a CUDA graph, then replayed
        fitter.fit(fw['vertices'], fw['joints'], num_iter=2)
    fitter.fit(fw['vertices'], fw['joints'], num_iter=1, initial_pose_rotvecs=p, initial_shape_betas=b)
    fitter.fit_with_known_pose(p, fw['vertices'], fw['joints'])
    fitter.fit_with_known_shape(b, fw['vertices'], fw['joints'], num_iter=2)
    BodyFitter(bm, enable_kid=True).to(dev).fit(fw['vertices'], fw['joints'], num_iter=1)


def forward_modes():
    for B in (1, 70, 200):
        p, b, t = params(B, J, S)
        o = bm(p, b, t)
        bm(rel_rotmats=torch.eye(3, device=dev).expand(B, J, 3, 3).contiguous(), shape_betas=b)
        bm(glob_rotmats=o['orientations'], trans=t, return_vertices=False)


def smplx_and_convert():
    bx = BodyModel('smplx_tiny').to(dev)
    fx = BodyFitter(bx).to(dev)
    p, b, t = params(45, bx.num_joints, bx.num_betas)
    fw = bx(p, b, t)
    fx.fit(fw['vertices'], fw['joints'], num_iter=2)
    import scipy.sparse as sp
    n_out, n_in = bx.num_vertices, V
    idx = rs.randint(0, n_in, size=(n_out, 3))
    w = rs.dirichlet([1, 1, 1], size=n_out).astype(np.float32)
    csr = sp.csr_matrix((w.reshape(-1), (np.repeat(np.arange(n_out), 3), idx.reshape(-1))), shape=(n_out, n_in))
    conv = BodyConverter(bm, bx, vertex_converter_csr=csr).to(dev)
    p, b, t = params(33, J, S)
    conv.convert(p, b, t, num_iter=1)


run('fit_ragged', fit_ragged)
run('fit_options', fit_options)
run('forward_modes', forward_modes)
run('smplx_and_convert', smplx_and_convert)
print('driver done')
