"""TEST INFRASTRUCTURE -- gradient fixtures from the UNMODIFIED reference (pt backend, CPU, torch autograd).

For a subset of the golden fit / forward cases (same inputs as ``tests/golden/<case>.npz``) this script draws seeded
cotangents for every output, back-propagates ``sum_k <cot_k, out_k>`` through the reference
(``smplfitter.pt.BodyFitter.fit`` / ``BodyModel.forward``) and stores the input gradients:

    tests/golden/grad_<case>.npz:  cot_<output>, ref_grad_<input>

``tests/test_adjoint_cpu.py`` holds the gradient evaluation of this repo (``smplfitter_b200/pt/_adjoint.py``) against
them on the CPU; ``tests/test_gpu_grad.py`` does the same through the CUDA ops' registered backward on the B200.

Run in the build container (needs /root/reference or oracle/_ref):  python -m oracle.make_grad_golden
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402
from oracle.make_golden import FIT_CASES, KNOWN_POSE_CASES, KNOWN_SHAPE_CASES, aux_call_kwargs  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

# fit cases whose options the gradient path supports, and the inputs differentiated in each
GRAD_FIT_CASES = {
    'fit_tiny_it1': ('target_vertices', 'target_joints'),
    'fit_tiny_it3': ('target_vertices', 'target_joints'),
    'fit_tiny_nojoints': ('target_vertices',),
    'fit_tiny_weights': ('target_vertices', 'target_joints', 'vertex_weights', 'joint_weights'),
    'fit_tiny_initial': ('target_vertices', 'target_joints', 'initial_pose_rotvecs', 'initial_shape_betas'),
    'fit_tiny_it4_noadjust': ('target_vertices', 'target_joints'),
    'fit_tiny_nojoints_vw': ('target_vertices', 'vertex_weights'),
    'fit_smplx_tiny_it3': ('target_vertices', 'target_joints'),
    'fit_tiny_kid': ('target_vertices', 'target_joints'),
    'fit_tiny_converter_style': ('target_vertices',),
    'fit_tiny_scale_target': ('target_vertices', 'target_joints'),
    'fit_tiny_scale_fit': ('target_vertices', 'target_joints'),
    'fit_tiny_share_beta': ('target_vertices', 'target_joints'),
    'fit_tiny_share_beta_scale': ('target_vertices', 'target_joints'),
}
GRAD_FORWARD_CASES = {'fwd_tiny': 'smpl_tiny', 'fwd_smplx_tiny': 'smplx_tiny'}
# fit_with_known_pose / fit_with_known_shape cases -> inputs differentiated (kshape_tiny_scale_fit is left out: the
# reference's scale_fit branch mixes instances and coordinates, see make_golden.py)
GRAD_KNOWN_POSE_CASES = {
    'kpose_tiny': ('pose_rotvecs', 'target_vertices', 'target_joints'),
    'kpose_tiny_weights': ('target_vertices', 'target_joints', 'vertex_weights', 'joint_weights'),
    'kpose_tiny_kid_nojoints': ('pose_rotvecs', 'target_vertices'),
    'kpose_tiny_scale_target': ('target_vertices', 'target_joints'),
    'kpose_tiny_share_beta': ('target_vertices', 'target_joints'),
}
GRAD_KNOWN_SHAPE_CASES = {
    'kshape_tiny': ('shape_betas', 'target_vertices', 'target_joints'),
    'kshape_tiny_nojoints': ('shape_betas', 'target_vertices'),
    'kshape_tiny_weights_init': ('target_vertices', 'target_joints', 'vertex_weights', 'joint_weights', 'initial_pose_rotvecs'),
}
FIT_OUTPUTS = ('shape_betas', 'trans', 'orientations', 'relative_orientations', 'pose_rotvecs', 'kid_factor', 'scale_corr')
FWD_OUTPUTS = ('joints', 'orientations', 'vertices')


def fit_inputs(name, g):
    """fit() kwargs of a golden fit case as float32 numpy arrays (tensor arguments) + plain options (the fitter's
    constructor options are ``FIT_CASES[name][2]``)."""
    _, _, _, _, _, _, fkw, flags = FIT_CASES[name]
    tens = dict(target_vertices=g['target_vertices'])
    if flags.get('joints'):
        tens['target_joints'] = g['target_joints']
    if 'in_vw' in g:
        tens['vertex_weights'] = g['in_vw']
    if 'in_jw' in g:
        tens['joint_weights'] = g['in_jw']
    if 'in_init_pose' in g:
        tens['initial_pose_rotvecs'] = g['in_init_pose']
        tens['initial_shape_betas'] = g['in_init_betas']
    return tens, dict(fkw)


def known_inputs(name, g):
    """(method, fitter kwargs, model name, tensor kwargs as numpy, plain options) of a known-pose / known-shape case."""
    if name in KNOWN_POSE_CASES:
        mname, fitkw, _, ckw, flags = KNOWN_POSE_CASES[name]
        kw = aux_call_kwargs(g, flags, ckw, lambda x: x)
        kw['pose_rotvecs'] = g['in_pose']
        method = 'fit_with_known_pose'
    else:
        mname, fitkw, _, ckw, flags = KNOWN_SHAPE_CASES[name]
        kw = aux_call_kwargs(g, flags, dict(ckw, requested_keys=['pose_rotvecs', 'relative_orientations']), lambda x: x)
        kw['shape_betas'] = g['in_betas']
        method = 'fit_with_known_shape'
    tens = {k: v for k, v in kw.items() if isinstance(v, np.ndarray)}
    opts = {k: v for k, v in kw.items() if not isinstance(v, np.ndarray)}
    return method, fitkw, mname, tens, opts


def cotangents(shapes: dict, seed: int):
    rs = np.random.RandomState(seed)
    return {k: rs.randn(*s).astype(np.float32) for k, s in shapes.items()}


def main():
    ref = refload.load()
    rpt = ref.pt
    for idx, (name, wrt) in enumerate(GRAD_FIT_CASES.items()):
        g = dict(np.load(os.path.join(GOLD, name + '.npz')))
        mname, mkw, fitkw = FIT_CASES[name][0], FIT_CASES[name][1], FIT_CASES[name][2]
        bm = rpt.BodyModel(mname, 'neutral', **mkw)
        fitter = rpt.BodyFitter(bm, **fitkw)
        tens, opts = fit_inputs(name, g)
        tt = {k: torch.from_numpy(v).clone().requires_grad_(k in wrt) for k, v in tens.items()}
        out = fitter.fit(**tt, **opts, requested_keys=['pose_rotvecs', 'shape_betas', 'relative_orientations'])
        cot = cotangents({k: tuple(out[k].shape) for k in FIT_OUTPUTS if k in out}, seed=900 + idx)
        sum((out[k] * torch.from_numpy(cot[k])).sum() for k in cot).backward()
        rec = {('cot_' + k): v for k, v in cot.items()}
        for k in wrt:
            gr = tt[k].grad.numpy()
            assert np.isfinite(gr).all() and np.abs(gr).max() > 0, (name, k)
            rec['ref_grad_' + k] = gr
        np.savez_compressed(os.path.join(GOLD, 'grad_' + name + '.npz'), **rec)
        print(f'[grad] {name}: ' + ' '.join(f"{k}={np.abs(rec['ref_grad_' + k]).max():.2e}" for k in wrt))
    for idx, (name, wrt) in enumerate({**GRAD_KNOWN_POSE_CASES, **GRAD_KNOWN_SHAPE_CASES}.items()):
        g = dict(np.load(os.path.join(GOLD, name + '.npz')))
        method, fitkw, mname, tens, opts = known_inputs(name, g)
        fitter = rpt.BodyFitter(rpt.BodyModel(mname, 'neutral'), **fitkw)
        tt = {k: torch.from_numpy(v).clone().requires_grad_(k in wrt) for k, v in tens.items()}
        out = getattr(fitter, method)(**tt, **opts)
        cot = cotangents({k: tuple(v.shape) for k, v in sorted(out.items())}, seed=970 + idx)
        sum((out[k] * torch.from_numpy(cot[k])).sum() for k in cot).backward()
        rec = {('cot_' + k): v for k, v in cot.items()}
        for k in wrt:
            gr = tt[k].grad.numpy()
            assert np.isfinite(gr).all() and np.abs(gr).max() > 0, (name, k)
            rec['ref_grad_' + k] = gr
        np.savez_compressed(os.path.join(GOLD, 'grad_' + name + '.npz'), **rec)
        print(f'[grad] {name}: ' + ' '.join(f"{k}={np.abs(rec['ref_grad_' + k]).max():.2e}" for k in wrt))
    for idx, (name, mname) in enumerate(GRAD_FORWARD_CASES.items()):
        g = dict(np.load(os.path.join(GOLD, name + '.npz')))
        bm = rpt.BodyModel(mname, 'neutral')
        tt = {k: torch.from_numpy(g[k]).clone().requires_grad_(True) for k in ('pose', 'betas', 'trans')}
        out = bm(tt['pose'], tt['betas'], tt['trans'])
        cot = cotangents({k: tuple(out[k].shape) for k in FWD_OUTPUTS}, seed=950 + idx)
        sum((out[k] * torch.from_numpy(cot[k])).sum() for k in FWD_OUTPUTS).backward()
        rec = {('cot_' + k): v for k, v in cot.items()}
        rec.update({('ref_grad_' + k): tt[k].grad.numpy() for k in tt})
        np.savez_compressed(os.path.join(GOLD, 'grad_' + name + '.npz'), **rec)
        print(f'[grad] {name}: ' + ' '.join(f"{k}={np.abs(tt[k].grad.numpy()).max():.2e}" for k in tt))


if __name__ == '__main__':
    main()
