"""CPU: host-side logic -- static masks bit-exact vs the reference fixtures, device tables
self-consistent, C-ABI library loads and exports every declared symbol (no compute calls)."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

from smplfitter_b200 import masks, modeldata
from tests import golden_cases as gc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('name', list(gc.MASK_CASES))
def test_masks_bit_exact(name):
    mname, mkw = gc.MASK_CASES[name]
    g = gc.load(name)
    data = modeldata.initialize(mname, **mkw)
    plan = masks.build_fit_plan(np.asarray(data.weights, np.float32), data.kintree_parents, mname)
    assert np.array_equal(plan.part_assignment, g['part_assignment'])
    assert np.array_equal(plan.used_vertex_indices, g['used_vertex_indices'])
    assert plan.multi_joint_parts == list(g['multi'])
    assert plan.bone_parts == list(g['bone'])
    assert plan.leaf_parts == list(g['leaf'])
    assert plan.adjustable_parts == list(g['adjustable'])
    for key in ('assemble_indices', 'bone_pairs', 'fk_js', 'fk_ps', 'adj_parts', 'adj_part_joints', 'cas_flat'):
        assert np.array_equal(getattr(plan, key), g[key]), key
    assert plan.fk_level_sizes == list(g['fk_level_sizes'])
    assert plan.adj_level_sizes == list(g['adj_level_sizes'])
    assert plan.cas_starts == list(g['cas_starts'])
    assert np.array_equal(plan.part_counts, g['part_counts'])
    assert np.array_equal(plan.center_matrix, g['center_matrix'])
    assert np.array_equal(plan.mjp_joint_membership, g['mjp_joint_membership'])
    assert np.array_equal(plan.part_matrix.sum(1), g['part_matrix_rowsum'])
    assert np.array_equal(plan.part_matrix.argmax(0), g['part_matrix_argmax'])


@pytest.mark.parametrize('mname', ['smpl_tiny', 'smplx_tiny', 'smpl'])
def test_device_tables(mname):
    from smplfitter_b200.pt import BodyFitter, BodyModel

    bm = BodyModel(mname)
    fitter = BodyFitter(bm)
    V, J = bm.num_vertices, bm.num_joints
    order = bm._t_order.numpy()
    inv = bm._t_inv_order.numpy()
    assert sorted(order.tolist()) == list(range(V))
    assert np.array_equal(inv[order], np.arange(V))
    part = fitter.part_assignment.numpy()
    seg_start, seg_part = bm._t_seg_start.numpy(), bm._t_seg_part.numpy()
    psb = bm._t_part_seg_begin.numpy()
    assert seg_start[0] == 0 and seg_start[-1] == V
    for s in range(len(seg_part)):
        vs = order[seg_start[s]:seg_start[s + 1]]
        assert 0 < len(vs) <= 64 and np.all(part[vs] == seg_part[s])
    for p in range(J):
        assert np.all(seg_part[psb[p]:psb[p + 1]] == p)
        n = sum(seg_start[s + 1] - seg_start[s] for s in range(psb[p], psb[p + 1]))
        assert n == int((part == p).sum())
    idx, w = bm._t_skin_idx.numpy(), bm._t_skin_w.numpy()
    dense = np.zeros((V, J), np.float32)
    np.add.at(dense, (np.arange(V)[:, None], idx), w)
    assert np.array_equal(dense, bm.weights.numpy())
    pf = bm._t_posedirs_fit.numpy()
    P = 9 * (J - 1)
    assert np.array_equal(pf[:, :P].reshape(V, 3, P), bm.posedirs.numpy()[order]) and not pf[:, P:].any()


def test_library_exports_declared_symbols():
    from smplfitter_b200 import _native

    lib = _native.lib()
    header = open(os.path.join(ROOT, 'include', 'smplfit_b200.h')).read()
    names = set(re.findall(r'\b(smplfit_[a-z_]+)\s*\(', header))
    assert {'smplfit_fit', 'smplfit_forward', 'smplfit_convert_vertices', 'smplfit_fit_known_pose'} <= names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.smplfit_struct_size(0) == ctypes.sizeof(_native.ModelStruct)
    assert lib.smplfit_struct_size(1) == ctypes.sizeof(_native.FitOpts)
    assert b'sm_100a' in lib.smplfit_version()


def test_no_cpu_fallback():
    """The product path refuses to run without CUDA instead of silently falling back."""
    import torch

    from smplfitter_b200.pt import BodyFitter, BodyModel

    bm = BodyModel('smpl_tiny')
    with pytest.raises(RuntimeError, match='CUDA'):
        bm(pose_rotvecs=torch.zeros(2, 72))
    with pytest.raises(RuntimeError, match='CUDA'):
        BodyFitter(bm).fit(torch.zeros(2, bm.num_vertices, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'smplfitter_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert 'oracle' not in src.replace('no oracle', ''), os.path.join(root, f)


def test_modules_are_scriptable():
    """torch.jit.script over model / fitter / converter compiles (reference: tests/conftest.py:38-39,
    pt/__init__.py:90 script the fitter); the scripted call still refuses to run without a GPU."""
    import warnings

    from smplfitter_b200.pt import BodyConverter, BodyFitter, BodyModel

    bm = BodyModel('smpl_tiny')
    fitter = BodyFitter(bm, enable_kid=True)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sm = torch.jit.script(bm)
        sf = torch.jit.script(fitter)
        sc = torch.jit.script(BodyConverter(bm, bm))
    assert 'requested_keys' in str(sf.fit.schema)
    assert hasattr(sc, 'convert_vertices')
    with pytest.raises(RuntimeError):
        sm(shape_betas=torch.zeros(1, 10))
    with pytest.raises(RuntimeError):
        sf.fit(torch.zeros(2, bm.num_vertices, 3))


@pytest.mark.parametrize('mname,kid', [('smpl_tiny', False), ('smplx_tiny', True)])
def test_closed_form_gram_tables(mname, kid):
    """The model constants behind k_gram_closed / k_shape_lite (include/smplfit_b200.h gcf_*, seg_slots, yj_*):
    a float64 emulation of the two kernels from the device tables reproduces the per-vertex normal equations
    G = sum_v jac^T jac, r = sum_v jac^T b of pt/bodyfitter.py:999-1048."""
    from scipy.spatial.transform import Rotation

    from smplfitter_b200.pt import BodyFitter, BodyModel
    from smplfitter_b200.pt.bodymodel import N_SLOTS

    bm = BodyModel(mname)
    f = BodyFitter(bm, enable_kid=kid)
    J, V, ns = bm.num_joints, bm.num_vertices, f._ns
    rng = np.random.default_rng(0)
    R = Rotation.random(J, random_state=1).as_matrix()
    T = rng.normal(size=(J, 3, 1 + ns))            # T_ext[k][c][0 | 1+s]
    t, vp = rng.normal(size=(V, 3)), rng.normal(size=(V, 3))  # internal vertex order
    rec = f._t_fit_rec.numpy().astype(np.float64)
    nsp = (ns + 1) // 2 * 2
    w4 = rec[:, 0:4]
    j4 = f._t_fit_rec.numpy()[:, 4:8].copy().view(np.int32)
    S = np.stack([rec[:, 8 + x * nsp:8 + x * nsp + ns] for x in range(3)], axis=1)  # (V,3,ns)
    # brute force
    Rb = np.einsum('vk,vkab->vab', w4, R[j4])
    Tb = np.einsum('vk,vkas->vas', w4, T[j4])
    jac = np.einsum('vab,vbs->vas', Rb, S) + Tb[:, :, 1:]
    b = t - (np.einsum('vab,vb->va', Rb, vp) + Tb[:, :, 0])
    G_ref, r_ref = np.einsum('vas,vat->st', jac, jac), np.einsum('vas,va->s', jac, b)
    # k_gram_closed
    iu, ju = np.triu_indices(ns)
    ng = len(iu)
    G = f._t_gcf_G0.numpy().copy()
    A = f._t_gcf_A.numpy().astype(np.float64)
    for p, (k, l) in enumerate(f._t_gcf_pairs.numpy()[:f._gcf_npairs]):
        G += np.einsum('ab,abe->e', R[k].T @ R[l], A[p, :, :ng].reshape(3, 3, ng))
    lstart, lk = f._t_gcf_lstart.numpy(), f._t_gcf_lk.numpy()
    Bm, Wh = f._t_gcf_Bm.numpy().astype(np.float64)[:, :, :ns], f._t_gcf_Wh.numpy().astype(np.float64)
    for l in range(J):
        Q = np.zeros((3, ns))
        for q in range(lstart[l], lstart[l + 1]):
            Q += R[lk[q]] @ Bm[q] + Wh[q] * T[lk[q]][:, 1:]
        Tl = T[l][:, 1:]
        G += np.einsum('ce,ce->e', Tl[:, ju], Q[:, iu]) + np.einsum('ce,ce->e', Tl[:, iu], Q[:, ju])
    Gm = np.zeros((ns, ns))
    Gm[iu, ju] = G
    Gm[ju, iu] = G
    assert np.abs(Gm - G_ref).max() <= 2e-6 * np.abs(G_ref).max()  # constants are stored in float32
    # k_shape_lite + k_lite_reduce
    seg_start, seg_slots = bm._t_seg_start.numpy(), bm._t_seg_slots.numpy()
    n_seg = len(seg_start) - 1
    z = np.einsum('vca,vc->va', Rb, b)
    r = np.einsum('vas,va->s', S, z)
    cells = np.zeros((n_seg * N_SLOTS, 3))
    for sgm in range(n_seg):
        lut = {int(j): sl for sl, j in enumerate(seg_slots[sgm]) if j >= 0}
        for i in range(seg_start[sgm], seg_start[sgm + 1]):
            for k in range(4):
                if w4[i, k] != 0:
                    cells[sgm * N_SLOTS + lut[int(j4[i, k])]] += w4[i, k] * b[i]
    yj_start, yj_entry = bm._t_yj_start.numpy(), bm._t_yj_entry.numpy()
    for k in range(J):
        Yk = cells[yj_entry[yj_start[k]:yj_start[k + 1]]].sum(axis=0)
        r += T[k][:, 1:].T @ Yk
    assert np.abs(r - r_ref).max() <= 1e-9 * np.abs(r_ref).max()


def test_body_flipper_host_logic():
    """BodyFlipper's host side (pt/bodyflipper.py:112-137): mirror assignment and the naive rotation-vector flip."""
    from smplfitter_b200.pt import BodyFlipper, BodyModel
    from smplfitter_b200.pt.bodyflipper import get_mirror_mapping, nearest_mirror_csr

    rs = np.random.RandomState(0)
    half = rs.randn(7, 3) + [2.0, 0, 0]
    pts = np.concatenate([half, half * [-1, 1, 1], [[0.0, 1.0, 2.0]]])  # 7 mirrored pairs + one point on the plane
    m = get_mirror_mapping(pts)
    assert np.array_equal(m, np.concatenate([np.arange(7, 14), np.arange(0, 7), [14]]))
    csr = nearest_mirror_csr(pts)
    assert np.allclose(csr @ pts * [-1, 1, 1], pts)  # mirror transfer + x flip maps a symmetric set onto itself
    bm = BodyModel('smpl_tiny')
    fl = BodyFlipper(bm)
    J = bm.num_joints
    rv = torch.from_numpy(rs.randn(5, 3 * J).astype(np.float32))
    got = fl.naive_flip_rotvecs(rv).numpy().reshape(5, J, 3)
    want = rv.numpy().reshape(5, J, 3)[:, fl.mirror_inds_joints.numpy()] * [1, -1, -1]
    assert np.array_equal(got, want.astype(np.float32))
    with pytest.raises(RuntimeError):
        fl.flip_vertices(torch.zeros(1, bm.num_vertices, 3))  # CPU module: no fallback


@pytest.mark.parametrize('mname', ['smpl_tiny', 'smplx_tiny'])
def test_slot_mask_and_gemm_operands(mname):
    """Host tables added for the TMA-staged kernels: fit_slot_mask replays the kernels' per-slot register cache
    (reset at every segment start, skipping zero weights), and gcf_AT_hi + gcf_AT_lo is the transposed, padded gcf_A."""
    from smplfitter_b200.pt import BodyFitter, BodyModel

    bm = BodyModel(mname)
    ft = BodyFitter(bm)
    rec = ft._t_fit_rec.numpy()
    w = rec[:, 0:4]
    idx = rec[:, 4:8].view(np.int32)
    seg = bm._t_seg_start.numpy()
    mask = ft._t_fit_slot_mask.numpy()
    assert mask.shape == (bm.num_vertices,) and mask.dtype == np.uint8
    for a, b in zip(seg[:-1], seg[1:]):
        # first vertex of a segment: every slot with a non-zero weight loads
        assert mask[a] == sum(1 << k for k in range(4) if w[a, k] != 0)
        for k in range(4):
            last = -1
            for i in range(a, b):
                need = w[i, k] != 0 and idx[i, k] != last
                assert bool(mask[i] >> k & 1) == bool(need), (i, k)
                if need:
                    last = idx[i, k]
    # the records past the last segment (unused vertices, if any) carry no bits that matter; segments tile [0, V)
    assert seg[0] == 0 and seg[-1] == bm.num_vertices
    A = ft._t_gcf_A.numpy()                      # (npairs, 9, NGP)
    AT = ft._t_gcf_AT_hi.numpy() + ft._t_gcf_AT_lo.numpy()
    ns = ft._ns
    ng = ns * (ns + 1) // 2
    assert AT.shape[0] % 256 == 0 and AT.shape[1] % 32 == 0
    assert np.array_equal(AT[:ng, :9 * A.shape[0]], A[:, :, :ng].reshape(-1, ng).T)
    assert not AT[ng:].any() and not AT[:, 9 * A.shape[0]:].any()
    hi = ft._t_gcf_AT_hi.numpy()
    assert not (hi.view(np.uint32) & np.uint32(0x1FFF)).any()  # tf32-exact high parts


def test_header_is_plain_c(tmp_path):
    """include/smplfit_b200.h is a C header (what a cgo / JNI / ctypes-free binding would include): it must compile
    as C11 on its own and declare every entry point the library exports."""
    import shutil
    import subprocess

    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    src = tmp_path / 'use_header.c'
    src.write_text(
        '#include "smplfit_b200.h"\n'
        'int probe(const smplfit_model_t* m, const smplfit_fit_opts_t* o) {\n'
        '  size_t a = smplfit_fit_workspace_bytes(m, 1, o, 1, 0, 0) + smplfit_forward_workspace_bytes(m, 1) +\n'
        '             smplfit_fit_host_workspace_bytes(m, 1, 1, o, 1) + smplfit_struct_size(0);\n'
        '  return (int)a + (smplfit_version() != 0) + (int)smplfit_launch_count(0);\n'
        '}\n')
    r = subprocess.run([gcc, '-std=c11', '-Wall', '-Werror', '-fsyntax-only', '-I', os.path.join(ROOT, 'include'), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_copies_get_their_own_handle():
    """A deep copy / unpickled copy of a module dispatches to ITS buffers (own registry handle), not to the original's."""
    import copy
    import pickle

    from smplfitter_b200.pt import BodyFitter, BodyModel, _ops

    bm = BodyModel('smpl_tiny')
    fitter = BodyFitter(bm)
    f2 = copy.deepcopy(fitter)
    assert f2._handle != fitter._handle and f2.body_model._handle != bm._handle
    assert _ops._get(f2._handle) is f2 and _ops._get(f2.body_model._handle) is f2.body_model
    bm3 = pickle.loads(pickle.dumps(bm))
    assert bm3._handle != bm._handle and _ops._get(bm3._handle) is bm3


def test_gradient_evaluation_is_reachable_only_from_backward_paths():
    """pt/_adjoint.py holds a differentiable torch evaluation of the fit.  It must never serve values: it is imported
    only inside the registered backward functions of the custom ops and inside the ``requires_grad`` branches of the
    known-pose / known-shape wrappers (whose values still come from their CUDA entry points)."""
    import ast
    import smplfitter_b200.pt as ptpkg

    pkg = os.path.dirname(ptpkg.__file__)
    allowed = {'_ops.py': {'_forward_backward', '_fit_backward'},
               'bodyfitter.py': {'fit_with_known_pose', 'fit_with_known_shape'}}
    for name in sorted(os.listdir(pkg)):
        if not name.endswith('.py') or name == '_adjoint.py':
            continue
        tree = ast.parse(open(os.path.join(pkg, name)).read())
        for fn in ast.walk(tree):
            if not isinstance(fn, (ast.FunctionDef, ast.Module)):
                continue
            for node in ast.iter_child_nodes(fn) if isinstance(fn, ast.Module) else ast.walk(fn):
                if isinstance(node, ast.ImportFrom) and any(a.name == '_adjoint' for a in node.names):
                    where = getattr(fn, 'name', '<module>')
                    assert where in allowed.get(name, set()), f'{name}: _adjoint imported in {where}'
    src = open(os.path.join(pkg, 'bodyfitter.py')).read()
    for chunk in src.split('from . import _adjoint')[:-1]:
        assert 'requires_grad' in chunk[-400:], 'the known-* wrappers may reach _adjoint only when a gradient is requested'


def test_fit_plan_on_random_trees():
    """masks.build_fit_plan against the oracle's independent re-derivation (OraclePlan) on random kinematic trees and
    skin weights, beyond the four models the reference-generated mask fixtures cover: partition, part kinds,
    children-and-self lists, used vertices, and the invariants the kernels rely on (levels cover every joint once,
    every parent precedes its children in the level order, the assembly order is a permutation)."""
    from hypothesis import given, settings, strategies as st

    from oracle import oracle_np
    from smplfitter_b200 import masks

    class Model:
        pass

    @settings(max_examples=40, deadline=None)
    @given(st.integers(3, 20), st.integers(30, 200), st.integers(0, 2 ** 31 - 1), st.booleans())
    def check(J, V, seed, smpl_family):
        rs = np.random.RandomState(seed)
        if smpl_family:
            J = max(J, 20)  # the SMPL-family adjustable list names joints up to 19
        parents = [0] + [int(rs.randint(0, i)) for i in range(1, J)]
        w = np.zeros((V, J), np.float32)
        for v in range(V):
            js = rs.choice(J, size=min(J, 1 + rs.randint(0, 4)), replace=False)
            w[v, js] = rs.dirichlet(np.ones(len(js))).astype(np.float32)
        name = 'smpl' if smpl_family else 'other'
        plan = masks.build_fit_plan(w, parents, name)
        m = Model()
        m.model_name, m.weights, m.parents, m.num_joints, m.num_vertices = name, w, parents, J, V
        ora = oracle_np.OraclePlan(m)
        assert np.array_equal(plan.part_assignment, ora.part)
        assert plan.children_and_self == ora.cas
        assert (plan.multi_joint_parts, plan.bone_parts, plan.leaf_parts) == (ora.multi, ora.bone, ora.leaf)
        assert plan.adjustable_parts == ora.adjustable
        assert np.array_equal(plan.used_vertex_indices, ora.used)
        assert sorted(plan.fk_js.tolist()) == list(range(1, J))
        seen = {0}
        for j, p in zip(plan.fk_js.tolist(), plan.fk_ps.tolist()):
            assert p == parents[j] and p in seen
            seen.add(j)
        order = plan.multi_joint_parts + plan.leaf_parts + plan.bone_parts
        assert sorted(order) == [i for i in range(J) if not (smpl_family and i in (10, 11))]
        assert sum(plan.fk_level_sizes) == J - 1 and sum(plan.adj_level_sizes) == len(plan.adj_parts)

    check()
