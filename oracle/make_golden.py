"""TEST INFRASTRUCTURE -- pins the oracle and writes the golden fixtures.

Runs the UNMODIFIED reference (``smplfitter.pt``, CPU, float32) from /root/reference/src
on the synthetic models (injected through ``smplfitter.common.initialize``, see
oracle/refload.py), compares ``oracle/oracle_np.py`` against it case by case, and writes
``tests/golden/*.npz`` (inputs + reference outputs) so the same checks run on the GPU box
where the reference tree does not exist.

    python -m oracle.make_golden            # check + (re)write fixtures

Only runs in the build container.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import oracle_np, refload  # noqa: E402
from smplfitter_b200 import modeldata  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

# name -> (model, model kwargs, fitter kwargs, B, pose scale, noise (m), fit kwargs, input flags)
FIT_CASES = {
    'fit_smpl_it3': ('smpl', {}, {}, 2, 0.1, 0.0,
                     dict(num_iter=3, beta_regularizer=1.0), dict(joints=True)),
    'fit_smpl_it3_noisy': ('smpl', {}, {}, 2, 0.3, 0.005,
                           dict(num_iter=3, beta_regularizer=1.0), dict(joints=True)),
    'fit_tiny_it1': ('smpl_tiny', {}, {}, 5, 0.3, 0.002,
                     dict(num_iter=1, beta_regularizer=1.0), dict(joints=True)),
    'fit_tiny_it3': ('smpl_tiny', {}, {}, 5, 0.3, 0.002,
                     dict(num_iter=3, beta_regularizer=1.0), dict(joints=True)),
    'fit_tiny_it4_noadjust': ('smpl_tiny', {}, {}, 4, 0.4, 0.002,
                              dict(num_iter=4, beta_regularizer=0.1, beta_regularizer2=0.05,
                                   final_adjust_rots=False), dict(joints=True)),
    'fit_tiny_weights': ('smpl_tiny', {}, {}, 4, 0.3, 0.003,
                         dict(num_iter=3, beta_regularizer=1.0), dict(joints=True, vw=True, jw=True)),
    'fit_tiny_vw_only': ('smpl_tiny', {}, {}, 4, 0.3, 0.003,
                         dict(num_iter=2, beta_regularizer=1.0), dict(joints=True, vw=True)),
    'fit_tiny_nojoints': ('smpl_tiny', {}, {}, 4, 0.3, 0.002,
                          dict(num_iter=3, beta_regularizer=1.0), dict(joints=False)),
    'fit_tiny_nojoints_vw': ('smpl_tiny', {}, {}, 4, 0.3, 0.002,
                             dict(num_iter=2, beta_regularizer=0.5), dict(joints=False, vw=True)),
    'fit_tiny_kid': ('smpl_tiny', {}, dict(enable_kid=True), 4, 0.3, 0.002,
                     dict(num_iter=2, beta_regularizer=1.0, kid_regularizer=10.0), dict(joints=True)),
    'fit_tiny_converter_style': ('smpl_tiny', {}, dict(enable_kid=True), 4, 0.3, 0.002,
                                 dict(num_iter=2, beta_regularizer=0.0, final_adjust_rots=False,
                                      kid_regularizer=1e9), dict(joints=False)),
    'fit_tiny_initial': ('smpl_tiny', {}, {}, 4, 0.3, 0.002,
                         dict(num_iter=2, beta_regularizer=1.0), dict(joints=True, initial=True)),
    'fit_tiny_scale_target': ('smpl_tiny', {}, {}, 4, 0.3, 0.002,
                              dict(num_iter=2, beta_regularizer=1.0, scale_target=True), dict(joints=True)),
    'fit_tiny_scale_fit': ('smpl_tiny', {}, {}, 4, 0.3, 0.002,
                           dict(num_iter=2, beta_regularizer=1.0, scale_fit=True), dict(joints=True)),
    'fit_tiny_share_beta': ('smpl_tiny', {}, {}, 5, 0.3, 0.002,
                            dict(num_iter=2, beta_regularizer=1.0, share_beta=True), dict(joints=True, same_betas=True)),
    'fit_smplx_tiny_it3': ('smplx_tiny', {}, {}, 3, 0.2, 0.002,
                           dict(num_iter=3, beta_regularizer=1.0), dict(joints=True)),
    'fit_smplx_tiny_nojoints': ('smplx_tiny', {}, {}, 3, 0.2, 0.002,
                                dict(num_iter=2, beta_regularizer=1.0), dict(joints=False)),
    'fit_smpl_subset1024': ('smpl', dict(vertex_subset_size=1024), {}, 3, 0.2, 0.002,
                            dict(num_iter=3, beta_regularizer=1.0), dict(joints=True)),
    'fit_smpl_betas6': ('smpl_tiny', dict(num_betas=6), {}, 3, 0.2, 0.002,
                        dict(num_iter=2, beta_regularizer=1.0), dict(joints=True)),
}
FORWARD_CASES = {'fwd_smpl': ('smpl', 3), 'fwd_tiny': ('smpl_tiny', 4), 'fwd_smplx_tiny': ('smplx_tiny', 3),
                 'fwd_smplx': ('smplx', 2)}
MASK_CASES = {'mask_smpl': ('smpl', {}), 'mask_smplx': ('smplx', {}), 'mask_tiny': ('smpl_tiny', {}),
              'mask_smplx_tiny': ('smplx_tiny', {}), 'mask_subset': ('smpl', dict(vertex_subset_size=1024))}


def case_inputs(name, model_name, mkw, B, pose_scale, noise, flags, seed):
    """Seeded on-manifold inputs: params -> reference forward -> + noise."""
    data = modeldata.initialize(model_name, **mkw)
    J, S, V = data.num_joints, data.shapedirs.shape[2], data.num_vertices
    rs = np.random.RandomState(seed)
    pose = (rs.randn(B, 3 * J) * pose_scale).astype(np.float32)
    betas = (rs.randn(B, S) * 0.5).astype(np.float32)
    if flags.get('same_betas'):
        betas = np.repeat(betas[:1], B, axis=0)
    trans = rs.randn(B, 3).astype(np.float32)
    inp = dict(pose=pose, betas=betas, trans=trans)
    if flags.get('vw'):
        inp['vw'] = rs.uniform(0.2, 1.5, size=(B, V)).astype(np.float32)
    if flags.get('jw'):
        inp['jw'] = rs.uniform(0.2, 1.5, size=(B, J)).astype(np.float32)
    if flags.get('initial'):
        inp['init_pose'] = (pose + rs.randn(B, 3 * J).astype(np.float32) * 0.05).astype(np.float32)
        inp['init_betas'] = (betas + rs.randn(B, S).astype(np.float32) * 0.2).astype(np.float32)
    inp['noise_v'] = (rs.randn(B, V, 3) * noise).astype(np.float32)
    inp['noise_j'] = (rs.randn(B, J, 3) * noise).astype(np.float32)
    return data, inp


def call_kwargs(inp, flags, fkw, tv, tj, conv):
    kw = dict(fkw)
    kw['target_vertices'] = conv(tv)
    if flags.get('joints'):
        kw['target_joints'] = conv(tj)
    if 'vw' in inp:
        kw['vertex_weights'] = conv(inp['vw'])
    if 'jw' in inp:
        kw['joint_weights'] = conv(inp['jw'])
    if 'init_pose' in inp:
        kw['initial_pose_rotvecs'] = conv(inp['init_pose'])
        kw['initial_shape_betas'] = conv(inp['init_betas'])
    kw['requested_keys'] = ['pose_rotvecs', 'shape_betas']
    return kw


def maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))) if np.size(a) else 0.0


def main():
    refload.load()
    import smplfitter.pt as rpt

    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    worst = {}

    # ---- masks -------------------------------------------------------------------
    for name, (mname, mkw) in MASK_CASES.items():
        bm = rpt.BodyModel(mname, 'neutral', **mkw)
        fr = rpt.BodyFitter(bm)
        data = modeldata.initialize(mname, **mkw)
        plan = oracle_np.OraclePlan(oracle_np.OracleModel(data, mname))
        pa = fr.part_assignment.numpy()
        assert np.array_equal(pa, plan.part), name
        assert np.array_equal(fr.used_vertex_indices.numpy(), plan.used), name
        assert fr.multi_joint_parts == plan.multi and fr.bone_parts == plan.bone and fr.leaf_parts == plan.leaf
        np.savez_compressed(
            os.path.join(GOLD, name + '.npz'),
            part_assignment=pa.astype(np.int16),
            used_vertex_indices=fr.used_vertex_indices.numpy().astype(np.int32),
            multi=np.array(fr.multi_joint_parts), bone=np.array(fr.bone_parts), leaf=np.array(fr.leaf_parts),
            adjustable=np.array(fr.adjustable_parts),
            assemble_indices=fr.assemble_indices.numpy(), bone_pairs=fr.bone_pairs.numpy(),
            fk_js=fr.fk_js.numpy(), fk_ps=fr.fk_ps.numpy(), fk_level_sizes=np.array(fr.fk_level_sizes),
            adj_parts=fr.adj_parts.numpy(), adj_level_sizes=np.array(fr.adj_level_sizes),
            adj_part_joints=fr.adj_part_joints.numpy(), cas_flat=fr.cas_flat.numpy(),
            cas_starts=np.array(fr.cas_starts), part_counts=fr.part_counts.numpy().reshape(-1),
            center_matrix=fr.center_matrix.numpy(), mjp_joint_membership=fr.mjp_joint_membership.numpy(),
            part_matrix_rowsum=fr.part_matrix.numpy().sum(1), part_matrix_argmax=fr.part_matrix.numpy().argmax(0).astype(np.int16),
            gram_supported=np.array(fr.gram_supported),
        )
        print(f'[mask] {name}: ok (V={bm.num_vertices}, used={len(plan.used)})')

    # ---- forward -----------------------------------------------------------------
    for name, (mname, B) in FORWARD_CASES.items():
        data, inp = case_inputs(name, mname, {}, B, 0.4, 0.0, {}, seed=100)
        bm = rpt.BodyModel(mname, 'neutral')
        om = oracle_np.OracleModel(data, mname)
        T = torch.from_numpy
        kid = np.linspace(-0.2, 0.3, B).astype(np.float32)
        ref = bm(T(inp['pose']), T(inp['betas']), T(inp['trans']), kid_factor=T(kid))
        ora = om.forward(inp['pose'], inp['betas'], inp['trans'], kid_factor=kid)
        d = max(maxdiff(ref[k].numpy(), ora[k]) for k in ('vertices', 'joints', 'orientations'))
        ref_g = bm(glob_rotmats=ref['orientations'], shape_betas=T(inp['betas'][:, :4]), trans=T(inp['trans']))
        ora_g = om.forward(glob_rotmats=ora['orientations'], shape_betas=inp['betas'][:, :4], trans=inp['trans'])
        d = max(d, maxdiff(ref_g['vertices'].numpy(), ora_g['vertices']))
        worst[name] = d
        assert d < 2e-6, (name, d)
        stride = max(1, data.num_vertices // 400)
        np.savez_compressed(
            os.path.join(GOLD, name + '.npz'), pose=inp['pose'], betas=inp['betas'], trans=inp['trans'], kid=kid,
            stride=np.array(stride), vertices=ref['vertices'].numpy()[:, ::stride], joints=ref['joints'].numpy(),
            orientations=ref['orientations'].numpy(), vertices_glob4=ref_g['vertices'].numpy()[:, ::stride],
            joints_glob4=ref_g['joints'].numpy(),
        )
        print(f'[fwd ] {name}: oracle-vs-reference max abs {d:.2e}')

    # ---- fit ---------------------------------------------------------------------
    for idx, (name, (mname, mkw, fitkw, B, ps, noise, fkw, flags)) in enumerate(FIT_CASES.items()):
        data, inp = case_inputs(name, mname, mkw, B, ps, noise, flags, seed=200 + idx)
        bm = rpt.BodyModel(mname, 'neutral', **mkw)
        fr = rpt.BodyFitter(bm, **fitkw)
        T = torch.from_numpy
        fw = bm(T(inp['pose']), T(inp['betas']), T(inp['trans']))
        tv = fw['vertices'].numpy() + inp['noise_v']
        tj = fw['joints'].numpy() + inp['noise_j']
        ref = fr.fit(**call_kwargs(inp, flags, fkw, tv, tj, T))
        ref = {k: v.numpy() for k, v in ref.items()}
        om = oracle_np.OracleModel(data, mname)
        of = oracle_np.OracleFitter(om, **fitkw)
        ora = of.fit(**call_kwargs(inp, flags, fkw, tv, tj, lambda x: x))
        # 'exact': the same algorithm in float64 on the same float32 inputs
        oracle_np.set_precision(np.float64)
        try:
            ofx = oracle_np.OracleFitter(oracle_np.OracleModel(data, mname), **fitkw)
            exact = ofx.fit(**call_kwargs(inp, flags, fkw, tv, tj, lambda x: x))
        finally:
            oracle_np.set_precision(np.float32)
        # reference self-noise: identical problem with the vertices renumbered (summation order)
        perm = np.random.RandomState(5).permutation(data.num_vertices)
        data_p = modeldata.apply_vertex_subset(data, perm)
        import smplfitter.common as rcommon
        saved_init = rcommon.initialize
        rcommon.initialize = lambda *a, **k: data_p
        try:
            fr_p = rpt.BodyFitter(rpt.BodyModel(mname, 'neutral'), **fitkw)
        finally:
            rcommon.initialize = saved_init
        inp_p = dict(inp)
        if 'vw' in inp:
            inp_p['vw'] = inp['vw'][:, perm]
        ref_p = fr_p.fit(**call_kwargs(inp_p, flags, fkw, tv[:, perm], tj, T))
        noise_o = np.abs(ref_p['orientations'].numpy() - ref['orientations']).max(axis=(0, 2, 3))
        noise_b = maxdiff(ref_p['shape_betas'].numpy(), ref['shape_betas'])
        diffs = {k: maxdiff(ref[k], ora[k]) for k in ref}
        dex = {k: maxdiff(ref[k], exact[k]) for k in ref}
        worst[name] = max(diffs[k] for k in ('shape_betas', 'trans'))
        print(f'[fit ] {name}: oracle-ref ' + ' '.join(f'{k}={v:.1e}' for k, v in diffs.items()))
        print(f'       ref-exact  ' + ' '.join(f'{k}={v:.1e}' for k, v in dex.items())
              + f' | ref self-noise orient={noise_o.max():.1e} betas={noise_b:.1e}')
        assert set(ref) == set(ora), (name, set(ref), set(ora))
        assert worst[name] < max(5e-5, 4 * noise_b), (name, diffs)
        # orientations: within the reference's own reproducibility on this input
        tol_o = np.maximum(1e-4, np.maximum(6 * noise_o, noise_o.max()))
        d_o = np.abs(ref['orientations'] - ora['orientations']).max(axis=(0, 2, 3))
        assert np.all(d_o <= tol_o), (name, d_o, tol_o)
        save = {('ref_' + k): v for k, v in ref.items()}
        save.update({('exact_' + k): np.asarray(v, np.float64) for k, v in exact.items()})
        save['ref_noise_orient'] = noise_o
        save['ref_noise_betas'] = np.array(noise_b)
        save.update(target_vertices=tv, target_joints=tj)
        for k in ('vw', 'jw', 'init_pose', 'init_betas', 'pose', 'betas', 'trans'):
            if k in inp:
                save['in_' + k] = inp[k]
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **save)

    print('worst oracle-vs-reference deviation per case:')
    for k, v in worst.items():
        print(f'  {k}: {v:.2e}')


if __name__ == '__main__':
    main()
