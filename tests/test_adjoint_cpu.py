"""Gradient path (smplfitter_b200/pt/_adjoint.py) on the CPU: the differentiable evaluation that the CUDA ops'
backward replays is held (a) against the float64 oracle values of the golden fit cases (it must be the same function),
(b) against gradients back-propagated through the UNMODIFIED reference (tests/golden/grad_*.npz, written by
oracle/make_grad_golden.py), (c) against central finite differences in float64, and (d) the hand-derived proj_so3
pull-back against torch.autograd.gradcheck, including a reflection case and equal singular values."""

import numpy as np
import pytest
import torch

from tests import golden_cases as gc
import smplfitter_b200.pt as pt
from oracle.make_golden import FIT_CASES
from oracle.make_grad_golden import FIT_OUTPUTS, FWD_OUTPUTS, GRAD_FIT_CASES, GRAD_FORWARD_CASES, fit_inputs
from smplfitter_b200.pt import _adjoint

def _model(name):
    mname, mkw = FIT_CASES[name][0], FIT_CASES[name][1]
    return pt.BodyModel(mname, **mkw)


def _run(bm, tens, opts, want_rel=True, name=None):
    """-> dict of the outputs present (FIT_OUTPUTS order)."""
    fitkw = FIT_CASES[name][2] if name else {}
    out = _adjoint.fit(bm, bm.num_betas, want_pose_rotvecs=True, want_rel_orient=want_rel, **tens, **opts, **fitkw)
    return {k: o for k, o in zip(FIT_OUTPUTS, out) if o is not None}


@pytest.mark.parametrize('name', sorted(GRAD_FIT_CASES))
def test_adjoint_evaluation_is_the_fit(name):
    """float64 evaluation == float64 oracle of the same case (1e-9), i.e. the function differentiated is the fit."""
    g = gc.load(name)
    tens, opts = fit_inputs(name, g)
    out = _run(_model(name), {k: torch.from_numpy(v).double() for k, v in tens.items()}, opts, name=name)
    assert set(out) == {k[6:] for k in g if k.startswith('exact_')}
    for k, o in out.items():
        # (1e9 kid regulariser of the converter-style case: the float64 solve itself is only good to ~1e-7 there)
        assert np.abs(o.numpy() - g['exact_' + k]).max() < (1e-6 if 'converter' in name else 1e-9), k


@pytest.mark.parametrize('dtype,tol', [(torch.float64, 2e-3), (torch.float32, 1e-2)])
@pytest.mark.parametrize('name', sorted(GRAD_FIT_CASES))
def test_fit_gradients_match_reference_autograd(name, dtype, tol):
    """The reference back-propagates in float32, so its gradient carries its own rounding; the float64 evaluation is
    expected within 2e-3 of it (relative to the gradient's scale), the float32 one within 1e-2 (x 5 on the SMPL-X
    stand-in, whose 55 small parts make the float32 gradient of either implementation that much noisier: the float64
    finite-difference test below is the exact check)."""
    if 'smplx' in name:
        tol *= 5
    g, gg = gc.load(name), gc.load('grad_' + name)
    tens, opts = fit_inputs(name, g)
    wrt = GRAD_FIT_CASES[name]
    tt = {k: torch.from_numpy(v).to(dtype).requires_grad_(k in wrt) for k, v in tens.items()}
    out = _run(_model(name), tt, opts, name=name)
    loss = sum((o * torch.from_numpy(gg['cot_' + k]).to(dtype)).sum() for k, o in out.items())
    grads = torch.autograd.grad(loss, [tt[k] for k in wrt])
    for k, gr in zip(wrt, grads):
        ref = gg['ref_grad_' + k]
        assert torch.isfinite(gr).all()
        err = np.abs(gr.numpy() - ref).max() / np.abs(ref).max()
        assert err < tol, (k, err)


@pytest.mark.parametrize('name', sorted(GRAD_FORWARD_CASES))
def test_forward_gradients_match_reference_autograd(name):
    g, gg = gc.load(name), gc.load('grad_' + name)
    bm = pt.BodyModel(GRAD_FORWARD_CASES[name])
    c = _adjoint.constants(bm, torch.float64, torch.device('cpu'))
    tt = {k: torch.from_numpy(g[k]).double().requires_grad_(True) for k in ('pose', 'betas', 'trans')}
    with torch.no_grad():  # the values of the fixture were taken with a kid factor
        out = _adjoint.lbs(c, tt['pose'], tt['betas'], tt['trans'], torch.from_numpy(g['kid']).double())
    for k, o in zip(FWD_OUTPUTS, out):
        o = o.numpy()[:, ::int(g['stride'])] if k == 'vertices' else o.numpy()  # (vertices are stored strided)
        assert np.abs(o - g[k]).max() < 2e-5, k
    out = _adjoint.lbs(c, tt['pose'], tt['betas'], tt['trans'])
    loss = sum((o * torch.from_numpy(gg['cot_' + k]).double()).sum() for k, o in zip(FWD_OUTPUTS, out))
    grads = torch.autograd.grad(loss, list(tt.values()))
    for k, gr in zip(tt, grads):
        ref = gg['ref_grad_' + k]
        assert np.abs(gr.numpy() - ref).max() / np.abs(ref).max() < 1e-4, k


@pytest.mark.parametrize('name', ['fit_tiny_it3', 'fit_tiny_weights', 'fit_tiny_nojoints', 'fit_tiny_kid',
                                  'fit_tiny_scale_fit', 'fit_tiny_share_beta', 'fit_tiny_share_beta_scale'])
def test_fit_gradient_vs_finite_differences(name):
    """Directional central difference in float64 (the reference's own check, tests/pt/test_fitter_grad.py:56-99, at
    float64 resolution instead of its 5 %)."""
    g, gg = gc.load(name), gc.load('grad_' + name)
    tens, opts = fit_inputs(name, g)
    bm = _model(name)
    base = {k: torch.from_numpy(v).double() for k, v in tens.items()}
    cot = {k[4:]: torch.from_numpy(v).double() for k, v in gg.items() if k.startswith('cot_')}

    def loss_of(tt):
        return sum((o * cot[k]).sum() for k, o in _run(bm, tt, opts, name=name).items())

    rs = np.random.RandomState(7)
    for key in GRAD_FIT_CASES[name]:
        tt = dict(base)
        tt[key] = base[key].clone().requires_grad_(True)
        gr, = torch.autograd.grad(loss_of(tt), tt[key])
        d = torch.from_numpy(rs.randn(*base[key].shape))
        d /= d.norm()
        eps = 1e-6
        with torch.no_grad():
            fd = (loss_of(dict(base, **{key: base[key] + eps * d})) - loss_of(dict(base, **{key: base[key] - eps * d}))) / (2 * eps)
        an = (gr * d).sum()
        assert abs(float(an - fd)) < 1e-6 * max(1.0, abs(float(fd))), (key, float(an), float(fd))


def test_proj_so3_pullback():
    torch.manual_seed(0)
    A = torch.randn(6, 3, 3, dtype=torch.float64)
    A[1] = -A[1] if torch.linalg.det(A[1]) > 0 else A[1]  # reflection branch
    Q = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64))[0]
    A[2] = Q @ torch.diag(torch.tensor([2.0, 2.0, 0.5], dtype=torch.float64)) @ Q.T  # equal singular values
    A[3] = 1.7 * Q  # all equal (isotropic covariance)
    A.requires_grad_(True)
    R = _adjoint.proj_so3(A)
    eye = torch.eye(3, dtype=torch.float64)
    assert (R @ R.transpose(-1, -2) - eye).abs().max() < 1e-12 and (torch.linalg.det(R) - 1).abs().max() < 1e-12
    assert torch.autograd.gradcheck(_adjoint.proj_so3, (A,), eps=1e-7, atol=1e-6)


def test_rotation_maps_differentiable_at_identity():
    rv = torch.zeros(2, 3, dtype=torch.float64, requires_grad=True)
    R = _adjoint.rotvec2mat(rv)
    back = _adjoint.mat2rotvec(R)
    g, = torch.autograd.grad(back.sum(), rv)
    assert torch.isfinite(g).all() and (g - 1).abs().max() < 1e-12
    rv2 = (torch.randn(5, 3, dtype=torch.float64) * 1.2).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda x: _adjoint.mat2rotvec(_adjoint.rotvec2mat(x)), (rv2,), eps=1e-7, atol=1e-6)
    # (like the reference's, the log map may return the angle > pi representative: compare as rotations)
    assert (_adjoint.rotvec2mat(_adjoint.mat2rotvec(_adjoint.rotvec2mat(rv2))) - _adjoint.rotvec2mat(rv2)).abs().max() < 1e-12


def test_backward_of_the_fit_op_slices_and_accumulates(monkeypatch):
    """The registered backward (pt/_ops.py) on the CPU: the op's CUDA body is replaced by the float32 evaluation, so
    this covers the wiring -- saved inputs, None / empty cotangents, batch slicing, broadcast (batch-1) inputs -- against
    a direct back-propagation through the same evaluation."""
    name = 'fit_tiny_initial'
    g = gc.load(name)
    tens, opts = fit_inputs(name, g)
    bm = _model(name)
    fitter = pt.BodyFitter(bm)

    def fake_impl(tv, tj, vw, jw, num_iter, reg, reg2, sreg, kreg, share, adj, st, sf, ip, ib, ik, keys):
        with torch.no_grad():
            o = _adjoint.fit(bm, bm.num_betas, tv, tj, vw, jw, num_iter, reg, reg2, adj, ip, ib,
                             'pose_rotvecs' in keys, 'relative_orientations' in keys, scale_regularizer=sreg,
                             kid_regularizer=kreg, share_beta=share, scale_target=st, scale_fit=sf, initial_kid_factor=ik)
        res = dict(zip(FIT_OUTPUTS, o))
        return {k: v for k, v in res.items() if v is not None}

    monkeypatch.setattr(fitter, '_fit_impl', fake_impl)
    monkeypatch.setattr(_adjoint, '_slices', lambda B, per, *a, **k: [(a, min(a + 3, B)) for a in range(0, B, 3)])
    tt = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in tens.items()}
    tt['initial_pose_rotvecs'] = tt['initial_pose_rotvecs'][:1].detach().clone().requires_grad_(True)  # broadcast input
    tt['initial_shape_betas'] = tt['initial_shape_betas'][:1].detach().clone().requires_grad_(True)
    out = fitter.fit(**tt, **opts, requested_keys=['pose_rotvecs'])
    assert out['pose_rotvecs'].requires_grad and out['shape_betas'].requires_grad
    loss = out['pose_rotvecs'].pow(2).sum() + out['shape_betas'].sum()  # trans / orientations get no cotangent
    loss.backward()
    t2 = {k: v.detach().clone().requires_grad_(True) for k, v in tt.items()}
    o2 = _adjoint.fit(bm, bm.num_betas, want_pose_rotvecs=True, **t2, **opts)
    (o2[4].pow(2).sum() + o2[0].sum()).backward()
    for k in tt:
        assert tt[k].grad is not None and torch.isfinite(tt[k].grad).all()
        scale = t2[k].grad.abs().max()
        assert (tt[k].grad - t2[k].grad).abs().max() < 1e-3 * scale, k
    # no gradient requested: the op's outputs do not track
    out = fitter.fit(**{k: v.detach() for k, v in tt.items()}, **opts)
    assert not out['pose_rotvecs'].requires_grad
    # scale estimation through the op: the extra output carries a gradient too
    for v in tt.values():
        v.grad = None
    out = fitter.fit(**tt, **opts, scale_target=True, requested_keys=['pose_rotvecs'])
    (out['scale_corr'].sum() + out['trans'].sum()).backward()
    assert all(torch.isfinite(v.grad).all() for v in tt.values()) and tt['target_vertices'].grad.abs().max() > 0


@pytest.mark.parametrize('post_translate', [True, False])
@pytest.mark.parametrize('kid', [False, True])
def test_rototranslate_moves_the_mesh_rigidly(post_translate, kid):
    """BodyModel.rototranslate (pt/bodymodel.py:382-453): the returned parameters pose the mesh at R x + t
    (post_translate) or R (x - t), checked with the torch evaluation of the forward pass."""
    bm = pt.BodyModel('smpl_tiny')
    c = _adjoint.constants(bm, torch.float64, torch.device('cpu'))
    torch.manual_seed(1)
    pose, betas, trans = torch.randn(72).double() * 0.3, torch.randn(10).double() * 0.5, torch.randn(3).double()
    kf = torch.tensor(0.4, dtype=torch.float64) if kid else None
    R = _adjoint.rotvec2mat(torch.tensor([0.3, -1.1, 0.6], dtype=torch.float64))
    t = torch.tensor([0.2, -0.5, 1.5], dtype=torch.float64)
    bm64 = pt.BodyModel('smpl_tiny').double()
    new_pose, new_trans = bm64.rototranslate(R, t, pose, betas, trans, kf, post_translate=post_translate)
    kb = None if kf is None else kf[None]
    v0 = _adjoint.lbs(c, pose[None], betas[None], trans[None], kb)[2][0]
    v1 = _adjoint.lbs(c, new_pose[None], betas[None], new_trans[None], kb)[2][0]
    want = v0 @ R.T + t if post_translate else (v0 - t) @ R.T
    assert (v1 - want).abs().max() < 1e-7
    with pytest.raises(ValueError):
        bm64.rototranslate(R, t, pose, None, trans)


# ---- fit_with_known_pose / fit_with_known_shape -------------------------------------------------------------------
from oracle.make_grad_golden import GRAD_KNOWN_POSE_CASES, GRAD_KNOWN_SHAPE_CASES, known_inputs  # noqa: E402

KNOWN_CASES = {**GRAD_KNOWN_POSE_CASES, **GRAD_KNOWN_SHAPE_CASES}


def _known_run(name, g, dtype, wrt=()):
    method, fitkw, mname, tens, opts = known_inputs(name, g)
    bm = pt.BodyModel(mname)
    tt = {k: torch.from_numpy(v).to(dtype).requires_grad_(k in wrt) for k, v in tens.items()}
    opts = dict(opts)
    if method == 'fit_with_known_pose':
        out = _adjoint.fit_with_known_pose(bm, bm.num_betas, bool(fitkw.get('enable_kid')), **tt, **opts)
    else:
        keys = opts.pop('requested_keys')
        out = _adjoint.fit_with_known_shape(bm, bm.num_betas, **tt, **opts, want_pose_rotvecs='pose_rotvecs' in keys,
                                            want_rel_orient='relative_orientations' in keys)
    return tt, out


@pytest.mark.parametrize('name', sorted(KNOWN_CASES) + ['kshape_tiny_scale_fit'])
def test_known_evaluations_are_the_methods(name):
    """Same result keys and values as the reference fixtures (float64 evaluation against the float32 reference, or
    against the float64 oracle where the fixture carries it)."""
    g = gc.load(name)
    _, out = _known_run(name, g, torch.float64)
    assert set(out) == {k[4:] for k in g if k.startswith('ref_') and k != 'ref_is_loose'}
    for k, o in out.items():
        if 'exact_' + k in g:
            # (the oracle normalises the float32 weights of the translation step in float32: 1e-8 in trans, which the
            # final adjustment of the small parts amplifies to 1e-5 in their rotations)
            tol = 1e-9 if 'weights' not in name else (1e-7 if k == 'trans' else 1e-4)
            assert np.abs(o.numpy() - g['exact_' + k]).max() < tol, k
        else:
            assert np.abs(o.numpy() - g['ref_' + k]).max() < (2e-3 if 'orient' in k else 5e-5), k


@pytest.mark.parametrize('name', sorted(KNOWN_CASES))
def test_known_gradients_match_reference_autograd(name):
    g, gg = gc.load(name), gc.load('grad_' + name)
    wrt = KNOWN_CASES[name]
    tt, out = _known_run(name, g, torch.float64, wrt)
    loss = sum((out[k[4:]] * torch.from_numpy(v).double()).sum() for k, v in gg.items() if k.startswith('cot_'))
    grads = torch.autograd.grad(loss, [tt[k] for k in wrt])
    for k, gr in zip(wrt, grads):
        ref = gg['ref_grad_' + k]
        err = np.abs(gr.numpy() - ref).max() / np.abs(ref).max()
        assert torch.isfinite(gr).all() and err < 3e-3, (k, err)


def test_differentiable_call_wrapper():
    """The wrapper of the known-* methods (values from one callable, gradient by sliced re-evaluation of another),
    with a broadcast (batch-1) input."""
    name = 'kshape_tiny'
    g = gc.load(name)
    method, fitkw, mname, tens, opts = known_inputs(name, g)
    bm = pt.BodyModel(mname)
    names = ['shape_betas', 'target_vertices', 'target_joints']
    tt = [torch.from_numpy(tens[k]).clone() for k in names]
    tt[0] = tt[0][:1]
    for x in tt:
        x.requires_grad_(True)
    run = lambda *xs: _adjoint.fit_with_known_shape(bm, bm.num_betas, *xs, num_iter=2)  # noqa: E731
    keys = ['trans', 'orientations', 'relative_orientations', 'pose_rotvecs']
    out = _adjoint.differentiable_call(run, run, keys, 4, torch.device('cpu'), 1e9, False, tt)  # 1 instance per slice
    (out['pose_rotvecs'].pow(2).sum() + out['trans'].sum()).backward()
    t2 = [x.detach().clone().requires_grad_(True) for x in tt]
    o2 = run(*t2)
    (o2['pose_rotvecs'].pow(2).sum() + o2['trans'].sum()).backward()
    for a, b in zip(tt, t2):
        assert (a.grad - b.grad).abs().max() < 1e-3 * b.grad.abs().max()


def test_convert_vertices_backward_is_the_transpose(monkeypatch):
    """The registered backward of the topology transfer against a dense M^T (CPU: the op's CUDA body replaced by a
    dense product)."""
    from oracle.make_golden import synthetic_converter_csr

    bm_in, bm_out = pt.BodyModel('smpl_tiny'), pt.BodyModel('smplx_tiny')
    csr = synthetic_converter_csr(bm_in.num_vertices, bm_out.num_vertices)
    conv = pt.BodyConverter(bm_in, bm_out, vertex_converter_csr=csr)
    M = torch.from_numpy(csr.toarray().astype(np.float32))
    monkeypatch.setattr(conv, '_convert_vertices_impl', lambda x: torch.einsum('oi,bic->boc', M, x))
    x = torch.randn(5, bm_in.num_vertices, 3, requires_grad=True)
    cot = torch.randn(5, bm_out.num_vertices, 3)
    (conv.convert_vertices(x) * cot).sum().backward()
    assert (x.grad - torch.einsum('oi,boc->bic', M, cot)).abs().max() < 1e-5


def test_converter_chain_gradient_in_float64():
    """forward -> transfer matrix -> converter-style fit (kid unknown pinned by a 1e9 regulariser, no joints), the
    chain BodyConverter.convert differentiates on the GPU: autograd against a central difference at eps = 1e-6."""
    from oracle.make_golden import synthetic_converter_csr

    bm_in, bm_out = pt.BodyModel('smpl_tiny'), pt.BodyModel('smplx_tiny')
    M = torch.from_numpy(synthetic_converter_csr(bm_in.num_vertices, bm_out.num_vertices).toarray()).double()
    ci = _adjoint.constants(bm_in, torch.float64, torch.device('cpu'))
    torch.manual_seed(8)
    pose, betas, trans = torch.randn(2, 72).double() * 0.2, torch.randn(2, 10).double() * 0.5, torch.randn(2, 3).double()
    cot = [torch.randn(2, 16).double(), torch.randn(2, 3).double(), torch.randn(2, 165).double()]

    def loss(p, b):
        verts = torch.einsum('oi,bic->boc', M, _adjoint.lbs(ci, p, b, trans)[2])
        o = _adjoint.fit(bm_out, 16, verts, num_iter=2, beta_regularizer=0.0, final_adjust_rots=False, enable_kid=True,
                         kid_regularizer=1e9)
        return (o[0] * cot[0]).sum() + (o[1] * cot[1]).sum() + (o[4] * cot[2]).sum()

    p, b = pose.clone().requires_grad_(True), betas.clone().requires_grad_(True)
    loss(p, b).backward()
    dp, db = torch.randn_like(pose), torch.randn_like(betas)
    dp, db = dp / dp.norm(), db / db.norm()
    with torch.no_grad():
        fd = (loss(pose + 1e-6 * dp, betas + 1e-6 * db) - loss(pose - 1e-6 * dp, betas - 1e-6 * db)).item() / 2e-6
    an = ((p.grad * dp).sum() + (b.grad * db).sum()).item()
    assert abs(an - fd) < 1e-5 * max(1.0, abs(fd)), (an, fd)


def test_two_restatements_agree_on_random_options():
    """The numpy oracle (per part / per joint loops) and the torch evaluation (vectorised, differentiable) are
    independent statements of the same algorithm: in float64 they must agree on random inputs and option mixes,
    beyond the golden cases."""
    from hypothesis import given, settings, strategies as st

    from oracle import oracle_np
    from smplfitter_b200 import modeldata

    data = modeldata.initialize('smpl_tiny')
    bm = pt.BodyModel('smpl_tiny')
    c = _adjoint.constants(bm, torch.float64, torch.device('cpu'))

    @settings(max_examples=12, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.integers(1, 3), st.booleans(), st.booleans(), st.booleans(), st.booleans(),
           st.sampled_from([0, 1, 2]))
    def check(seed, num_iter, joints, weights, adjust, share, scale_mode):
        rs = np.random.RandomState(seed)
        B = 3
        pose, betas, trans = rs.randn(B, 72) * 0.25, rs.randn(B, 10) * 0.6, rs.randn(B, 3)
        with torch.no_grad():
            j, _, v = _adjoint.lbs(c, *(torch.from_numpy(x) for x in (pose, betas, trans)))
        tv = v.numpy() * (1.05 if scale_mode else 1.0) + rs.randn(B, bm.num_vertices, 3) * 0.003
        tj = j.numpy() * (1.05 if scale_mode else 1.0) + rs.randn(B, 24, 3) * 0.003 if joints else None
        vw = rs.uniform(0.3, 1.5, (B, bm.num_vertices)) if weights else None
        jw = rs.uniform(0.3, 1.5, (B, 24)) if (weights and joints) else None
        kw = dict(num_iter=num_iter, beta_regularizer=float(rs.uniform(0, 2)), beta_regularizer2=float(rs.uniform(0, 0.5)),
                  final_adjust_rots=adjust, share_beta=share, scale_target=scale_mode == 1, scale_fit=scale_mode == 2,
                  scale_regularizer=float(rs.uniform(0, 1)))
        oracle_np.set_precision(np.float64)
        try:
            ora = oracle_np.OracleFitter(oracle_np.OracleModel(data, 'smpl_tiny')).fit(
                tv, tj, vw, jw, requested_keys=['pose_rotvecs', 'shape_betas', 'relative_orientations'], **kw)
        finally:
            oracle_np.set_precision(np.float32)
        T = lambda x: None if x is None else torch.from_numpy(x)  # noqa: E731
        out = _adjoint.fit(bm, 10, T(tv), T(tj), T(vw), T(jw), want_pose_rotvecs=True, want_rel_orient=True, **kw)
        for k, o in zip(FIT_OUTPUTS, out):
            if o is not None:
                assert np.abs(o.numpy() - ora[k]).max() < 1e-8, (k, kw)

    check()
