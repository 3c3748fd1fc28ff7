"""CPU: the numpy oracle reproduces the reference outputs stored in tests/golden/."""

import numpy as np
import pytest

from oracle import oracle_np
from smplfitter_b200 import modeldata
from tests import golden_cases as gc


@pytest.mark.parametrize('name', list(gc.MASK_CASES))
def test_oracle_masks(name):
    mname, mkw = gc.MASK_CASES[name]
    g = gc.load(name)
    plan = oracle_np.OraclePlan(oracle_np.OracleModel(modeldata.initialize(mname, **mkw), mname))
    assert np.array_equal(plan.part, g['part_assignment'])
    assert np.array_equal(plan.used, g['used_vertex_indices'])
    assert plan.multi == list(g['multi']) and plan.bone == list(g['bone']) and plan.leaf == list(g['leaf'])


@pytest.mark.parametrize('name', list(gc.FORWARD_CASES))
def test_oracle_forward(name):
    mname, _ = gc.FORWARD_CASES[name]
    g = gc.load(name)
    om = oracle_np.OracleModel(modeldata.initialize(mname), mname)
    out = om.forward(g['pose'], g['betas'], g['trans'], kid_factor=g['kid'])
    s = int(g['stride'])
    assert np.abs(out['vertices'][:, ::s] - g['vertices']).max() < 5e-6
    assert np.abs(out['joints'] - g['joints']).max() < 5e-6
    assert np.abs(out['orientations'] - g['orientations']).max() < 5e-6
    out4 = om.forward(glob_rotmats=g['orientations'], shape_betas=g['betas'][:, :4], trans=g['trans'])
    assert np.abs(out4['vertices'][:, ::s] - g['vertices_glob4']).max() < 5e-6
    assert np.abs(out4['joints'] - g['joints_glob4']).max() < 5e-6


@pytest.mark.parametrize('name', list(gc.FIT_CASES))
def test_oracle_fit(name):
    mname, mkw, fitkw = gc.FIT_CASES[name][:3]
    g = gc.load(name)
    om = oracle_np.OracleModel(modeldata.initialize(mname, **mkw), mname)
    out = oracle_np.OracleFitter(om, **fitkw).fit(**gc.fit_call_kwargs(name, g))
    assert {('ref_' + k) for k in out} == {k for k in g if k.startswith('ref_') and not k.startswith('ref_noise')}
    tol_b = max(5e-5, 4 * float(g['ref_noise_betas']))
    assert np.abs(out['shape_betas'] - g['ref_shape_betas']).max() < tol_b
    assert np.abs(out['trans'] - g['ref_trans']).max() < tol_b
    d = np.abs(out['orientations'] - g['ref_orientations']).max(axis=(0, 2, 3))
    assert np.all(d <= gc.orient_tolerance(g)), (d, gc.orient_tolerance(g))
    for k in ('kid_factor', 'scale_corr'):
        if k in out:
            assert np.abs(out[k] - g['ref_' + k]).max() < 5e-5


@pytest.mark.parametrize('name', list(gc.KNOWN_POSE_CASES))
def test_oracle_known_pose(name):
    """fit_with_known_pose of the unmodified reference (pt/bodyfitter.py:552-653)."""
    mname, fitkw, _, ckw, flags = gc.KNOWN_POSE_CASES[name]
    g = gc.load(name)
    of = oracle_np.OracleFitter(oracle_np.OracleModel(modeldata.initialize(mname), mname), **fitkw)
    out = of.fit_with_known_pose(g['in_pose'], **gc.aux_call_kwargs(g, flags, ckw, lambda x: x))
    assert {('ref_' + k) for k in out} == {k for k in g if k.startswith('ref_')}
    for k in out:
        assert np.abs(out[k] - g['ref_' + k]).max() < 5e-5, k


@pytest.mark.parametrize('name', list(gc.KNOWN_SHAPE_CASES))
def test_oracle_known_shape(name):
    """fit_with_known_shape of the unmodified reference (pt/bodyfitter.py:656-838)."""
    mname, fitkw, _, ckw, flags = gc.KNOWN_SHAPE_CASES[name]
    g = gc.load(name)
    of = oracle_np.OracleFitter(oracle_np.OracleModel(modeldata.initialize(mname), mname), **fitkw)
    kw = gc.aux_call_kwargs(g, flags, dict(ckw, requested_keys=['pose_rotvecs', 'relative_orientations']), lambda x: x)
    out = oracle_np.fit_with_known_shape(of, g['in_betas'], **kw)
    loose = bool(g['ref_is_loose'])  # see oracle/make_golden.py: the reference's scale_fit broadcasting
    assert {('ref_' + k) for k in out} == {k for k in g if k.startswith('ref_') and k != 'ref_is_loose'}
    assert np.abs(out['trans'] - g['ref_trans']).max() < (2e-4 if loose else 5e-5)
    assert np.abs(out['orientations'] - g['ref_orientations']).max() < (5e-3 if loose else 2e-3)
    if 'scale_corr' in out:
        assert np.abs(out['scale_corr'] - g['ref_scale_corr']).max() < 5e-5


@pytest.mark.parametrize('name', list(gc.CONVERT_CASES))
def test_oracle_convert(name):
    """BodyConverter.convert of the unmodified reference (pt/bodyconverter.py:48-127) = forward + CSR + fit."""
    m_in, m_out, _, ckw, branch = gc.CONVERT_CASES[name]
    g = gc.load(name)
    om_in = oracle_np.OracleModel(modeldata.initialize(m_in), m_in)
    om_out = oracle_np.OracleModel(modeldata.initialize(m_out), m_out)
    verts = om_in.forward(g['in_pose'], g['in_betas'], g['in_trans'])['vertices']
    if 'csr_data' in g:
        verts = oracle_np.convert_vertices_csr(g['csr_indptr'], g['csr_indices'], g['csr_data'], verts)
    assert np.abs(verts - g['ref_converted_vertices']).max() < 5e-6
    of = oracle_np.OracleFitter(om_out, enable_kid=True)
    if branch == 'known_shape':
        out = oracle_np.fit_with_known_shape(of, g['in_known_betas'], verts, num_iter=ckw.get('num_iter', 1),
                                             final_adjust_rots=False, requested_keys=['pose_rotvecs'])
        keys = ('pose_rotvecs', 'trans')
    elif branch == 'known_pose':
        out = of.fit_with_known_pose(g['in_known_pose'], verts, beta_regularizer=0.0, kid_regularizer=1e9)
        keys = ('shape_betas', 'trans')
    else:
        out = of.fit(verts, num_iter=ckw.get('num_iter', 1), beta_regularizer=0.0, final_adjust_rots=False,
                     kid_regularizer=1e9, requested_keys=['pose_rotvecs', 'shape_betas'])
        keys = ('pose_rotvecs', 'shape_betas', 'trans')
    for k in keys:
        assert np.abs(out[k] - g['ref_' + k]).max() < (5e-3 if k == 'pose_rotvecs' else 1e-4), k
