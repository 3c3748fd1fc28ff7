"""Gradients through the CUDA ops (reference feature, /root/reference/tests/pt/test_fitter_grad.py:31-99).

The values come from the CUDA kernels, the gradient from the backward registered on the ``smplfit_b200::fit`` /
``::forward`` ops (pt/_adjoint.py).  Checked here: the reference's own two gradient tests restated on this API
(finite + non-zero for smpl / smplx and num_iter 1 / 3; directional finite differences of the CUDA fit within 5 %),
the gradients back-propagated through the UNMODIFIED reference for the golden cases (tests/golden/grad_*.npz), and the
batch slicing of the backward on a batch that does not fit one slice."""

import numpy as np
import pytest
import torch

from tests import golden_cases as gc
import smplfitter_b200.pt as pt
from oracle.make_golden import FIT_CASES
from oracle.make_grad_golden import FWD_OUTPUTS, GRAD_FIT_CASES, GRAD_FORWARD_CASES, fit_inputs
from smplfitter_b200.pt import _adjoint

pytestmark = pytest.mark.gpu


def _targets(model_name, batch_size, seed=0, pose_scale=0.1):
    torch.manual_seed(seed)
    bm = pt.BodyModel(model_name, num_betas=10).cuda()
    pose = (torch.randn(batch_size, bm.num_joints * 3) * pose_scale).cuda()
    shape = (torch.randn(batch_size, 10) * 0.5).cuda()
    trans = torch.randn(batch_size, 3).cuda()
    with torch.no_grad():
        out = bm(pose_rotvecs=pose, shape_betas=shape, trans=trans)
    return bm, out['vertices'].detach(), out['joints'].detach()


def _loss(fit):
    return sum(fit[k].pow(2).sum() for k in ['pose_rotvecs', 'shape_betas', 'trans'])


@pytest.mark.parametrize('model_name', ['smpl', 'smplx'])
@pytest.mark.parametrize('num_iter', [1, 3])
def test_fitter_grad_finite(model_name, num_iter):
    bm, target_v, target_j = _targets(model_name, 2)
    fitter = pt.BodyFitter(bm).cuda()
    tv = target_v.clone().requires_grad_(True)
    tj = target_j.clone().requires_grad_(True)
    fit = fitter.fit(target_vertices=tv, target_joints=tj, num_iter=num_iter, beta_regularizer=1.0,
                     requested_keys=['pose_rotvecs', 'shape_betas', 'trans'])
    _loss(fit).backward()
    for name, g in [('target_vertices', tv.grad), ('target_joints', tj.grad)]:
        assert g is not None, name
        assert g.is_cuda and torch.isfinite(g).all(), name
        assert g.abs().max().item() > 0, name


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_fitter_grad_vs_finite_diff(seed):
    """Directional derivative from the registered backward against a central difference of the CUDA fit itself."""
    bm, target_v, target_j = _targets('smpl', 1, seed)
    fitter = pt.BodyFitter(bm).cuda()

    def loss_fn(tv, tj):
        return _loss(fitter.fit(target_vertices=tv, target_joints=tj, num_iter=1, beta_regularizer=1.0,
                                requested_keys=['pose_rotvecs', 'shape_betas', 'trans']))

    tv = target_v.clone().requires_grad_(True)
    tj = target_j.clone().requires_grad_(True)
    loss_fn(tv, tj).backward()
    g = torch.Generator().manual_seed(seed + 100)
    dv = torch.randn(target_v.shape, generator=g)
    dj = torch.randn(target_j.shape, generator=g)
    dv, dj = (dv / dv.norm()).cuda(), (dj / dj.norm()).cuda()
    ag = (tv.grad * dv).sum().item() + (tj.grad * dj).sum().item()
    eps = 1e-2
    with torch.no_grad():
        lp = loss_fn(target_v + eps * dv, target_j + eps * dj).item()
        lm = loss_fn(target_v - eps * dv, target_j - eps * dj).item()
    fd = (lp - lm) / (2 * eps)
    rel = abs(ag - fd) / max(abs(ag), abs(fd), 1e-3)
    assert rel < 5e-2, f'seed={seed}: autograd={ag:.4e}, fd={fd:.4e}, rel={rel:.3e}'


@pytest.mark.parametrize('name', sorted(GRAD_FIT_CASES))
def test_fit_gradients_match_reference_autograd(name):
    """Every option family of fit(): joints or not, weights, initial guesses, kid factor, scale estimation, shared betas."""
    g, gg = gc.load(name), gc.load('grad_' + name)
    tens, opts = fit_inputs(name, g)
    wrt = GRAD_FIT_CASES[name]
    mname, mkw, fitkw = FIT_CASES[name][0], FIT_CASES[name][1], FIT_CASES[name][2]
    fitter = pt.BodyFitter(pt.BodyModel(mname, **mkw).cuda(), **fitkw).cuda()
    tt = {k: torch.from_numpy(v).cuda().requires_grad_(k in wrt) for k, v in tens.items()}
    out = fitter.fit(**tt, **opts, requested_keys=['pose_rotvecs', 'shape_betas', 'relative_orientations'])
    cot = {k[4:]: torch.from_numpy(v).cuda() for k, v in gg.items() if k.startswith('cot_')}
    assert set(cot) <= set(out)
    sum((out[k] * c).sum() for k, c in cot.items()).backward()
    tol = 5e-2 if 'smplx' in name else 1e-2
    for k in wrt:
        ref = gg['ref_grad_' + k]
        gr = tt[k].grad.cpu().numpy()
        assert np.isfinite(gr).all(), k
        err = np.abs(gr - ref).max() / np.abs(ref).max()
        assert err < tol, (k, err)


@pytest.mark.parametrize('name', sorted(GRAD_FORWARD_CASES))
def test_forward_gradients_match_reference_autograd(name):
    g, gg = gc.load(name), gc.load('grad_' + name)
    bm = pt.BodyModel(GRAD_FORWARD_CASES[name]).cuda()
    tt = {k: torch.from_numpy(g[k]).cuda().requires_grad_(True) for k in ('pose', 'betas', 'trans')}
    out = bm(tt['pose'], tt['betas'], tt['trans'])
    sum((out[k] * torch.from_numpy(gg['cot_' + k]).cuda()).sum() for k in FWD_OUTPUTS).backward()
    for k in tt:
        ref = gg['ref_grad_' + k]
        assert np.abs(tt[k].grad.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-4, k


def test_forward_gradient_other_rotation_inputs():
    """rel_rotmats / glob_rotmats / kid_factor inputs and return_vertices=False: finite differences of the CUDA forward."""
    bm = pt.BodyModel('smpl_tiny').cuda()
    torch.manual_seed(3)
    B, J = 3, bm.num_joints
    rel = _adjoint.rotvec2mat(torch.randn(B, J, 3) * 0.3).cuda()
    betas, kid = (torch.randn(B, 10) * 0.5).cuda(), (torch.rand(B) * 0.5).cuda()
    cot = torch.randn(B, bm.num_vertices, 3).cuda()
    for key in ('rel_rotmats', 'glob_rotmats'):
        rot = rel if key == 'rel_rotmats' else bm(rel_rotmats=rel, return_vertices=False)['orientations']
        x = rot.clone().requires_grad_(True)
        k = kid.clone().requires_grad_(True)
        (bm(shape_betas=betas, kid_factor=k, **{key: x})['vertices'] * cot).sum().backward()
        d = torch.randn_like(rot)
        d /= d.norm()
        eps = 1e-2
        with torch.no_grad():
            f = lambda r, kk: (bm(shape_betas=betas, kid_factor=kk, **{key: r})['vertices'] * cot).sum().item()  # noqa: E731
            fd = (f(rot + eps * d, kid) - f(rot - eps * d, kid)) / (2 * eps)
            fdk = (f(rot, kid + eps) - f(rot, kid - eps)) / (2 * eps)
        an = (x.grad * d).sum().item()
        assert abs(an - fd) < 2e-2 * max(abs(fd), 1.0), (key, an, fd)
        assert abs(k.grad.sum().item() - fdk) < 2e-2 * max(abs(fdk), 1.0), (key, k.grad.sum().item(), fdk)
    p = (torch.randn(B, 3 * J) * 0.2).cuda().requires_grad_(True)
    out = bm(pose_rotvecs=p, return_vertices=False)
    assert 'vertices' not in out
    out['joints'].sum().backward()
    assert torch.isfinite(p.grad).all() and p.grad.abs().max() > 0


def test_backward_slices_large_batch(monkeypatch):
    """A batch that needs several slices gives, per instance, the gradient of that instance fitted alone."""
    bm, tv, tj = _targets('smpl', 70, seed=5)
    fitter = pt.BodyFitter(bm).cuda()
    calls = []
    real = _adjoint._slices
    monkeypatch.setattr(_adjoint, '_slices', lambda B, per, *a, **k: calls.append(B) or real(B, per, 16 * per))
    a = tv.clone().requires_grad_(True)
    _loss(fitter.fit(a, tj, num_iter=2, requested_keys=['pose_rotvecs', 'shape_betas', 'trans'])).backward()
    assert calls == [70]
    for i in (0, 37, 69):
        b = tv[i:i + 1].clone().requires_grad_(True)
        _loss(fitter.fit(b, tj[i:i + 1], num_iter=2, requested_keys=['pose_rotvecs', 'shape_betas', 'trans'])).backward()
        assert (a.grad[i] - b.grad[0]).abs().max() < 2e-3 * b.grad.abs().max(), i


def test_gradient_refinement_lowers_the_loss():
    """What the reference uses the gradients for (pt/bodyfitter_opt.py:131-255): starting from the closed-form fit, a few
    Adam steps on pose / shape / translation through the CUDA forward reduce the vertex error to a noisy target."""
    bm, tv, tj = _targets('smpl', 8, seed=9)
    tv = tv + 0.01 * torch.randn_like(tv)
    fitter = pt.BodyFitter(bm).cuda()
    fit = fitter.fit(tv, tj, num_iter=1, requested_keys=['pose_rotvecs', 'shape_betas', 'trans'])
    params = [fit[k].detach().clone().requires_grad_(True) for k in ('pose_rotvecs', 'shape_betas', 'trans')]
    opt = torch.optim.Adam(params, lr=2e-3)
    losses = []
    for _ in range(15):
        opt.zero_grad()
        loss = (bm(*params)['vertices'] - tv).pow(2).sum(-1).mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


def test_body_fitter_opt():
    """pt/bodyfitter_opt.py: refine_steps=0 is the closed-form fit; the refinement lowers the vertex error of a
    one-iteration fit without final adjustment on noisy targets (no shape regulariser: with the default weight of 1 the
    objective trades vertex error for smaller betas, in the reference as here)."""
    from smplfitter_b200.pt.bodyfitter_opt import BodyFitterOpt, rot6d_to_rotmat, rotmat_to_rot6d

    bm, tv, tj = _targets('smpl', 6, seed=11, pose_scale=0.4)
    tv = tv + 0.005 * torch.randn_like(tv)
    opt = BodyFitterOpt(bm).cuda()
    plain = opt.fit(tv, tj, num_iter=1)
    ref = opt.fitter.fit(tv, tj, num_iter=1, requested_keys=['pose_rotvecs', 'shape_betas', 'trans'])
    for k in ('pose_rotvecs', 'shape_betas', 'trans'):
        assert torch.equal(plain[k], ref[k]), k
    err = lambda r: (bm(r['pose_rotvecs'], r['shape_betas'], r['trans'])['vertices'] - tv).norm(dim=-1).mean().item()  # noqa: E731
    start = opt.fitter.fit(tv, tj, num_iter=1, beta_regularizer=0.0, final_adjust_rots=False,
                           requested_keys=['pose_rotvecs', 'shape_betas', 'trans'])
    refined = opt.fit(tv, tj, num_iter=1, beta_regularizer=0.0, refine_steps=30, refine_lr=0.003)
    assert set(refined) == {'pose_rotvecs', 'shape_betas', 'trans'}
    assert err(refined) < 0.97 * err(start), (err(refined), err(start))
    R = _adjoint.rotvec2mat(torch.randn(5, 3))
    assert (rot6d_to_rotmat(rotmat_to_rot6d(R)) - R).abs().max() < 1e-5


def test_share_beta_backward_is_one_slice(monkeypatch):
    """share_beta couples the instances: the backward must not cut the batch."""
    bm, tv, tj = _targets('smpl_tiny', 6, seed=4)
    fitter = pt.BodyFitter(bm).cuda()
    monkeypatch.setattr(_adjoint, '_slices', lambda *a, **k: pytest.fail('share_beta backward was sliced'))
    a = tv.clone().requires_grad_(True)
    out = fitter.fit(a, tj, num_iter=2, share_beta=True, requested_keys=['pose_rotvecs', 'shape_betas', 'trans'])
    assert (out['shape_betas'] - out['shape_betas'][:1]).abs().max() == 0
    _loss(out).backward()
    assert torch.isfinite(a.grad).all() and a.grad.abs().max() > 0


# ---- fit_with_known_pose / fit_with_known_shape: values from their CUDA entry points, gradient through the wrapper ----
from oracle.make_grad_golden import GRAD_KNOWN_POSE_CASES, GRAD_KNOWN_SHAPE_CASES, known_inputs  # noqa: E402


@pytest.mark.parametrize('name', sorted({**GRAD_KNOWN_POSE_CASES, **GRAD_KNOWN_SHAPE_CASES}))
def test_known_pose_shape_gradients_match_reference_autograd(name):
    g, gg = gc.load(name), gc.load('grad_' + name)
    wrt = {**GRAD_KNOWN_POSE_CASES, **GRAD_KNOWN_SHAPE_CASES}[name]
    method, fitkw, mname, tens, opts = known_inputs(name, g)
    fitter = pt.BodyFitter(pt.BodyModel(mname).cuda(), **fitkw).cuda()
    tt = {k: torch.from_numpy(v).cuda().requires_grad_(k in wrt) for k, v in tens.items()}
    out = getattr(fitter, method)(**tt, **opts)
    with torch.no_grad():
        plain = getattr(fitter, method)(**{k: v.detach() for k, v in tt.items()}, **opts)
    assert set(out) == set(plain)
    for k in out:  # the values are the CUDA path's
        assert torch.equal(out[k], plain[k]) and out[k].requires_grad, k
    sum((out[k[4:]] * torch.from_numpy(v).cuda()).sum() for k, v in gg.items() if k.startswith('cot_')).backward()
    for k in wrt:
        ref = gg['ref_grad_' + k]
        gr = tt[k].grad.cpu().numpy()
        err = np.abs(gr - ref).max() / np.abs(ref).max()
        assert np.isfinite(gr).all() and err < 1e-2, (k, err)


def test_scripted_fitter_and_foreign_inputs_carry_gradients():
    """The backward is registered on the custom op, so it is there for the TorchScript-compiled fitter too; inputs in
    float64 or living on the CPU get their gradient back in their own dtype / on their own device."""
    bm, tv, tj = _targets('smpl_tiny', 3, seed=2)
    fitter = pt.BodyFitter(bm).cuda()
    scripted = torch.jit.script(fitter)
    keys = ['pose_rotvecs', 'shape_betas', 'trans']
    a = tv.clone().requires_grad_(True)
    _loss(fitter.fit(a, tj, num_iter=2, requested_keys=keys)).backward()
    b = tv.clone().requires_grad_(True)
    _loss(scripted.fit(b, tj, num_iter=2, requested_keys=keys)).backward()
    # (the re-evaluation sums with atomics on the GPU: equal up to the summation order)
    assert (a.grad - b.grad).abs().max() < 1e-4 * a.grad.abs().max()
    c = tv.double().cpu().requires_grad_(True)
    _loss(fitter.fit(c, tj, num_iter=2, requested_keys=keys)).backward()
    assert c.grad.dtype == torch.float64 and not c.grad.is_cuda
    assert (c.grad.float().cuda() - a.grad).abs().max() < 1e-4 * a.grad.abs().max()
    with torch.no_grad():
        out = fitter.fit(a, tj, num_iter=2, requested_keys=keys)
    assert not out['trans'].requires_grad


def test_inference_never_evaluates_the_adjoint(monkeypatch):
    """Without a requires_grad input every entry point is the CUDA path alone, in grad mode too."""
    def boom(*a, **k):
        raise AssertionError('pt/_adjoint.py evaluated on the inference path')

    for name in ('fit', 'lbs', 'fit_with_known_pose', 'fit_with_known_shape', '_pullback', 'differentiable_call'):
        monkeypatch.setattr(_adjoint, name, boom)
    bm, tv, tj = _targets('smpl_tiny', 3, seed=6)
    fitter = pt.BodyFitter(bm).cuda()
    res = fitter.fit(tv, tj, num_iter=2, requested_keys=['pose_rotvecs', 'shape_betas', 'trans'])
    assert not res['trans'].requires_grad
    out = bm(res['pose_rotvecs'], res['shape_betas'], res['trans'])
    assert not out['vertices'].requires_grad
    fitter.fit_with_known_pose(res['pose_rotvecs'], tv, tj)
    fitter.fit_with_known_shape(res['shape_betas'], tv, tj, num_iter=2)
    pt.BodyConverter(bm, bm).cuda().convert(res['pose_rotvecs'], res['shape_betas'], res['trans'])


def test_converter_is_differentiable_end_to_end():
    """BodyConverter.convert = forward op -> topology transfer op -> fit op: a directional finite difference of the
    converted parameters with respect to the input pose / betas against the chained backward.  The transfer matrix is
    a smoothing of the same topology (0.8 v_i + 0.2 v_{i+1}), so the converted mesh is a body and the fit is smooth at
    the finite-difference step; with the random barycentric stand-in of the golden cases the fitted finger rotations
    jump within eps = 1e-2 (the chained gradient still matches at eps = 1e-6 in float64, checked on the CPU)."""
    import scipy.sparse as sp

    bm_in, bm_out = pt.BodyModel('smpl_tiny').cuda(), pt.BodyModel('smpl_tiny').cuda()
    V, J = bm_in.num_vertices, bm_in.num_joints
    csr = (sp.identity(V) * 0.8 + sp.csr_matrix((np.full(V, 0.2), (np.arange(V), (np.arange(V) + 1) % V)), shape=(V, V))).tocsr()
    conv = pt.BodyConverter(bm_in, bm_out, vertex_converter_csr=csr).cuda()
    assert conv.has_converter
    torch.manual_seed(8)
    pose, betas, trans = (torch.randn(2, 3 * J) * 0.2).cuda(), (torch.randn(2, 10) * 0.5).cuda(), torch.randn(2, 3).cuda()
    cot = {k: torch.randn(s).cuda() for k, s in (('pose_rotvecs', (2, 3 * J)), ('shape_betas', (2, 10)), ('trans', (2, 3)))}

    def loss(p, b):
        out = conv.convert(p, b, trans, num_iter=2)
        return sum((out[k] * cot[k]).sum() for k in cot)

    p, b = pose.clone().requires_grad_(True), betas.clone().requires_grad_(True)
    loss(p, b).backward()
    assert torch.isfinite(p.grad).all() and torch.isfinite(b.grad).all() and p.grad.abs().max() > 0
    dp, db = torch.randn_like(pose), torch.randn_like(betas)
    dp, db = dp / dp.norm(), db / db.norm()
    eps = 1e-2
    with torch.no_grad():
        fd = (loss(pose + eps * dp, betas + eps * db) - loss(pose - eps * dp, betas - eps * db)).item() / (2 * eps)
    an = ((p.grad * dp).sum() + (b.grad * db).sum()).item()
    assert abs(an - fd) < 5e-2 * max(abs(an), abs(fd), 1e-3), (an, fd)
