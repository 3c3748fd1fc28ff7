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
