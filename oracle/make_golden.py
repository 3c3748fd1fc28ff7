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
    # BASELINE.json configs[2]: full-size SMPL-X shape (10475 vertices, 55 joints, 16 betas)
    'fit_smplx_it3': ('smplx', {}, {}, 2, 0.15, 0.002,
                      dict(num_iter=3, beta_regularizer=1.0), dict(joints=True)),
    # (appended last: the case index seeds the inputs of every case)
    'fit_tiny_share_beta_scale': ('smpl_tiny', {}, {}, 5, 0.3, 0.002,
                                  dict(num_iter=2, beta_regularizer=1.0, share_beta=True, scale_target=True,
                                       scale_regularizer=0.5), dict(joints=True, same_betas=True)),
}
FORWARD_CASES = {'fwd_smpl': ('smpl', 3), 'fwd_tiny': ('smpl_tiny', 4), 'fwd_smplx_tiny': ('smplx_tiny', 3),
                 'fwd_smplx': ('smplx', 2)}
# fit_with_known_pose / fit_with_known_shape / BodyConverter.convert of the unmodified reference
# name -> (model, fitter kwargs, B, call kwargs, input flags)
KNOWN_POSE_CASES = {
    'kpose_tiny': ('smpl_tiny', {}, 4, dict(beta_regularizer=0.5), dict(joints=True)),
    'kpose_tiny_weights': ('smpl_tiny', {}, 4, dict(beta_regularizer=1.0, beta_regularizer2=0.1),
                           dict(joints=True, vw=True, jw=True)),
    'kpose_tiny_kid_nojoints': ('smpl_tiny', dict(enable_kid=True), 3,
                                dict(beta_regularizer=0.0, kid_regularizer=1e9), dict(joints=False)),
    'kpose_tiny_scale_target': ('smpl_tiny', {}, 3, dict(beta_regularizer=1.0, scale_target=True), dict(joints=True)),
    'kpose_tiny_share_beta': ('smpl_tiny', {}, 4, dict(beta_regularizer=1.0, share_beta=True), dict(joints=True, same_betas=True)),
}
KNOWN_SHAPE_CASES = {
    'kshape_tiny': ('smpl_tiny', {}, 4, dict(num_iter=2, final_adjust_rots=True), dict(joints=True)),
    'kshape_tiny_nojoints': ('smpl_tiny', dict(enable_kid=True), 3, dict(num_iter=1, final_adjust_rots=False),
                             dict(joints=False)),
    'kshape_tiny_weights_init': ('smpl_tiny', {}, 4, dict(num_iter=2, final_adjust_rots=True),
                                 dict(joints=True, vw=True, jw=True, init_pose=True)),
    # the reference's scale_fit branch multiplies a (B,) scale with (B,3) / (B,V,3) arrays (pt/bodyfitter.py:1676,
    # :779-780): it only runs for batch == 3 and then applies instance c's scale to coordinate c.  Noise-free targets
    # with one common scale make that mix-up harmless, so the fixture pins the intended per-instance computation.
    'kshape_tiny_scale_fit': ('smpl_tiny', {}, 3, dict(num_iter=2, final_adjust_rots=True, scale_fit=True),
                              dict(joints=True, scale=1.07, noise=0.0)),
}
# name -> (model in, model out, B, convert kwargs, branch)
CONVERT_CASES = {
    'convert_tiny_default': ('smpl_tiny', 'smplx_tiny', 3, dict(num_iter=1), 'default'),
    'convert_tiny_it2': ('smpl_tiny', 'smplx_tiny', 3, dict(num_iter=2), 'default'),
    'convert_tiny_known_pose': ('smpl_tiny', 'smplx_tiny', 3, dict(), 'known_pose'),
    'convert_tiny_known_shape': ('smpl_tiny', 'smplx_tiny', 3, dict(num_iter=2), 'known_shape'),
    'convert_tiny_same_topology': ('smpl_tiny', 'smpl_tiny', 3, dict(num_iter=1), 'default'),
}
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
    if flags.get('same_betas'):
        inp['betas'][:] = inp['betas'][:1]
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


def synthetic_converter_csr(v_in, v_out, seed=8):
    """Barycentric-style transfer matrix (3 non-negative weights per row summing to one) standing in for the
    licensed smpl2smplx_deftrafo_setup.pkl (common.py:425-429)."""
    import scipy.sparse as sp

    rs = np.random.RandomState(seed)
    cols = rs.randint(0, v_in, size=(v_out, 3))
    w = rs.dirichlet([1, 1, 1], size=v_out).astype(np.float32)
    return sp.csr_matrix((w.reshape(-1), (np.repeat(np.arange(v_out), 3), cols.reshape(-1))), shape=(v_out, v_in))


def aux_inputs(mname, B, flags, seed):
    """Seeded inputs of the known-pose / known-shape cases: on-manifold targets + 2 mm noise."""
    data = modeldata.initialize(mname)
    J, S, V = data.num_joints, data.shapedirs.shape[2], data.num_vertices
    rs = np.random.RandomState(seed)
    inp = dict(pose=(rs.randn(B, 3 * J) * 0.3).astype(np.float32), betas=(rs.randn(B, S) * 0.7).astype(np.float32),
               trans=rs.randn(B, 3).astype(np.float32))
    if flags.get('same_betas'):
        inp['betas'][:] = inp['betas'][:1]
    if flags.get('vw'):
        inp['vw'] = rs.uniform(0.2, 1.5, size=(B, V)).astype(np.float32)
    if flags.get('jw'):
        inp['jw'] = rs.uniform(0.2, 1.5, size=(B, J)).astype(np.float32)
    if flags.get('init_pose'):
        inp['init_pose'] = (inp['pose'] + rs.randn(B, 3 * J).astype(np.float32) * 0.05).astype(np.float32)
    inp['noise_v'] = (rs.randn(B, V, 3) * flags.get('noise', 0.002)).astype(np.float32)
    inp['noise_j'] = (rs.randn(B, J, 3) * flags.get('noise', 0.002)).astype(np.float32)
    return data, inp


def aux_call_kwargs(g, flags, ckw, conv):
    """kwargs shared by fit_with_known_pose / fit_with_known_shape from a fixture dict."""
    kw = dict(ckw)
    kw['target_vertices'] = conv(g['target_vertices'])
    if flags.get('joints'):
        kw['target_joints'] = conv(g['target_joints'])
    if 'in_vw' in g:
        kw['vertex_weights'] = conv(g['in_vw'])
    if 'in_jw' in g:
        kw['joint_weights'] = conv(g['in_jw'])
    if 'in_init_pose' in g:
        kw['initial_pose_rotvecs'] = conv(g['in_init_pose'])
    return kw


def gen_aux(rpt, worst):
    """Known-pose / known-shape / converter fixtures from the unmodified reference; pins the oracle's versions."""
    T = torch.from_numpy
    ident = lambda x: x  # noqa: E731
    for idx, (name, (mname, fitkw, B, ckw, flags)) in enumerate(KNOWN_POSE_CASES.items()):
        data, inp = aux_inputs(mname, B, flags, seed=400 + idx)
        bm = rpt.BodyModel(mname, 'neutral')
        fr = rpt.BodyFitter(bm, **fitkw)
        fw = bm(T(inp['pose']), T(inp['betas']), T(inp['trans']))
        g = dict(target_vertices=fw['vertices'].numpy() + inp['noise_v'], target_joints=fw['joints'].numpy() + inp['noise_j'],
                 in_pose=inp['pose'])
        for k in ('vw', 'jw'):
            if k in inp:
                g['in_' + k] = inp[k]
        ref = fr.fit_with_known_pose(pose_rotvecs=T(inp['pose']), **aux_call_kwargs(g, flags, ckw, T))
        ref = {k: v.numpy() for k, v in ref.items()}
        of = oracle_np.OracleFitter(oracle_np.OracleModel(data, mname), **fitkw)
        ora = of.fit_with_known_pose(inp['pose'], **aux_call_kwargs(g, flags, ckw, ident))
        d = {k: maxdiff(ref[k], ora[k]) for k in ref}
        worst[name] = max(d.values())
        print(f'[kpos] {name}: oracle-ref ' + ' '.join(f'{k}={v:.1e}' for k, v in d.items()))
        assert set(ref) == set(ora), (name, set(ref), set(ora))
        assert worst[name] < 5e-5, (name, d)
        g.update({('ref_' + k): v for k, v in ref.items()})
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **g)

    for idx, (name, (mname, fitkw, B, ckw, flags)) in enumerate(KNOWN_SHAPE_CASES.items()):
        data, inp = aux_inputs(mname, B, flags, seed=500 + idx)
        bm = rpt.BodyModel(mname, 'neutral')
        fr = rpt.BodyFitter(bm, **fitkw)
        fw = bm(T(inp['pose']), T(inp['betas']), T(inp['trans']))
        sc = np.float32(flags.get('scale', 1.0))
        g = dict(target_vertices=sc * fw['vertices'].numpy() + inp['noise_v'],
                 target_joints=sc * fw['joints'].numpy() + inp['noise_j'], in_betas=inp['betas'])
        for k in ('vw', 'jw', 'init_pose'):
            if k in inp:
                g['in_' + k] = inp[k]
        ckw2 = dict(ckw, requested_keys=['pose_rotvecs', 'relative_orientations'])
        ref = fr.fit_with_known_shape(shape_betas=T(inp['betas']), **aux_call_kwargs(g, flags, ckw2, T))
        ref = {k: v.numpy() for k, v in ref.items()}
        of = oracle_np.OracleFitter(oracle_np.OracleModel(data, mname), **fitkw)
        ora = oracle_np.fit_with_known_shape(of, inp['betas'], **aux_call_kwargs(g, flags, ckw2, ident))
        d = {k: maxdiff(ref[k], ora[k]) for k in ref}
        worst[name] = max(d[k] for k in d if k in ('trans', 'scale_corr'))
        print(f'[kshp] {name}: oracle-ref ' + ' '.join(f'{k}={v:.1e}' for k, v in d.items()))
        assert set(ref) == set(ora), (name, set(ref), set(ora))
        # (scale_fit: what is left of the reference's scale mix-up once the three scales agree to ~1e-4)
        loose = bool(ckw.get('scale_fit'))
        assert worst[name] < (2e-4 if loose else 5e-5) and d['orientations'] < (5e-3 if loose else 2e-3), (name, d)
        # float64 evaluation of the same algorithm: the yardstick for the rotation outputs
        oracle_np.set_precision(np.float64)
        try:
            ofx = oracle_np.OracleFitter(oracle_np.OracleModel(data, mname), **fitkw)
            exact = oracle_np.fit_with_known_shape(ofx, inp['betas'], **aux_call_kwargs(g, flags, ckw2, ident))
        finally:
            oracle_np.set_precision(np.float32)
        g.update({('ref_' + k): v for k, v in ref.items()})
        g.update({('exact_' + k): np.asarray(v, np.float64) for k, v in exact.items()})
        g['ref_is_loose'] = np.array(loose)
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **g)

    for idx, (name, (m_in, m_out, B, ckw, branch)) in enumerate(CONVERT_CASES.items()):
        d_in, d_out = modeldata.initialize(m_in), modeldata.initialize(m_out)
        rs = np.random.RandomState(600 + idx)
        J_in, J_out = d_in.num_joints, d_out.num_joints
        pose = (rs.randn(B, 3 * J_in) * 0.25).astype(np.float32)
        betas = (rs.randn(B, d_in.shapedirs.shape[2]) * 0.6).astype(np.float32)
        trans = rs.randn(B, 3).astype(np.float32)
        bm_in, bm_out = rpt.BodyModel(m_in, 'neutral'), rpt.BodyModel(m_out, 'neutral')
        conv = rpt.BodyConverter(bm_in, bm_out)
        g = dict(in_pose=pose, in_betas=betas, in_trans=trans)
        csr = None
        if d_in.num_vertices != d_out.num_vertices:
            csr = synthetic_converter_csr(d_in.num_vertices, d_out.num_vertices, seed=8 + idx)
            conv.vertex_converter_csr = torch.sparse_csr_tensor(
                torch.from_numpy(csr.indptr.astype(np.int64)), torch.from_numpy(csr.indices.astype(np.int64)),
                torch.from_numpy(csr.data), size=csr.shape)
            g.update(csr_indptr=csr.indptr.astype(np.int32), csr_indices=csr.indices.astype(np.int32), csr_data=csr.data)
        kw = dict(ckw)
        if branch == 'known_pose':
            # a plausible output pose: the input pose on the shared body joints, zero elsewhere
            kp = np.zeros((B, 3 * J_out), np.float32)
            n = 3 * min(J_in, J_out, 22)
            kp[:, :n] = pose[:, :n]
            g['in_known_pose'] = kp
            kw['known_output_pose_rotvecs'] = T(kp)
        if branch == 'known_shape':
            kb = (rs.randn(B, d_out.shapedirs.shape[2]) * 0.3).astype(np.float32)
            g['in_known_betas'] = kb
            kw['known_output_shape_betas'] = T(kb)
        ref = conv.convert(T(pose), T(betas), T(trans), **kw)
        ref = {k: v.numpy() for k, v in ref.items()}
        g['ref_converted_vertices'] = conv.convert_vertices(bm_in(T(pose), T(betas), T(trans))['vertices']).contiguous().numpy()
        # the oracle's composition of the same three steps
        om_in, om_out = oracle_np.OracleModel(d_in, m_in), oracle_np.OracleModel(d_out, m_out)
        verts = om_in.forward(pose, betas, trans)['vertices']
        if csr is not None:
            verts = oracle_np.convert_vertices_csr(csr.indptr, csr.indices, csr.data, verts)
        assert maxdiff(verts, g['ref_converted_vertices']) < 2e-6, name
        of = oracle_np.OracleFitter(om_out, enable_kid=True)
        if branch == 'known_shape':
            ora = oracle_np.fit_with_known_shape(of, g['in_known_betas'], verts, num_iter=kw.get('num_iter', 1),
                                                 final_adjust_rots=False, requested_keys=['pose_rotvecs'])
            ora = dict(pose_rotvecs=ora['pose_rotvecs'], trans=ora['trans'])
        elif branch == 'known_pose':
            ora = of.fit_with_known_pose(g['in_known_pose'], verts, beta_regularizer=0.0, kid_regularizer=1e9)
            ora = dict(shape_betas=ora['shape_betas'], trans=ora['trans'])
        else:
            ora = of.fit(verts, num_iter=kw.get('num_iter', 1), beta_regularizer=0.0, final_adjust_rots=False,
                         kid_regularizer=1e9, requested_keys=['pose_rotvecs', 'shape_betas'])
            ora = dict(pose_rotvecs=ora['pose_rotvecs'], shape_betas=ora['shape_betas'], trans=ora['trans'])
        d = {k: maxdiff(ref[k], ora[k]) for k in ref}
        worst[name] = max(d[k] for k in d if k != 'pose_rotvecs')
        print(f'[conv] {name}: oracle-ref ' + ' '.join(f'{k}={v:.1e}' for k, v in d.items()))
        assert set(ref) == set(ora), (name, set(ref), set(ora))
        assert worst[name] < 1e-4 and d.get('pose_rotvecs', 0.0) < 5e-3, (name, d)
        g.update({('ref_' + k): v for k, v in ref.items()})
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **g)


def maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))) if np.size(a) else 0.0


def main():
    refload.load()
    import smplfitter.pt as rpt

    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    worst = {}
    only = sys.argv[sys.argv.index('--only') + 1] if '--only' in sys.argv else None
    if only == 'aux':
        gen_aux(rpt, worst)
        return

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
        if only is not None and only != name:
            continue
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

    gen_aux(rpt, worst)

    print('worst oracle-vs-reference deviation per case:')
    for k, v in worst.items():
        print(f'  {k}: {v:.2e}')


if __name__ == '__main__':
    main()
