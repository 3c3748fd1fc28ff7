"""TEST INFRASTRUCTURE -- CPU (numpy) restatement of the reference's pt hot path.

This is the parity oracle for the CUDA kernels: a from-scratch numpy statement of the
*algorithm* of ``smplfitter.pt`` (BodyModel.forward, BodyFitter.fit and its stages,
BodyConverter.convert_vertices) written per body part / per joint, i.e. the way a CUDA
thread sees it, not as the reference's one-hot GEMM formulation.  Each function cites the
reference lines it follows (paths relative to /root/reference/src/smplfitter/).

PINNING: ``oracle/make_golden.py`` runs the unmodified reference (pt backend, CPU) on the
synthetic models and (a) checks this file against it, (b) writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` re-checks this file against those fixtures everywhere.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  Nothing under ``smplfitter_b200/`` does.
"""

from __future__ import annotations

import numpy as np

F32 = np.float32  # working precision; ``set_precision(np.float64)`` turns this file into the exact evaluator


def set_precision(dtype):
    """Switch the working dtype (float32 = mirrors the reference, float64 = 'exact' arithmetic
    of the same algorithm, used to measure how far the reference's own fp32 rounding is from
    the true solution of each stage)."""
    global F32
    F32 = dtype


# --------------------------------------------------------------------------------------
# rotation helpers (pt/rotation.py)
# --------------------------------------------------------------------------------------
def divide_no_nan(a, b):
    """pt/rotation.py:8-11."""
    safe = np.where(b == 0, np.ones_like(b), b)
    return np.where(b == 0, np.zeros_like(a / safe), a / safe)


def rotvec2mat(rv):
    """Rodrigues, element formulas of pt/rotation.py:236-258."""
    rv = np.asarray(rv, dtype=F32)
    angle = np.linalg.norm(rv, axis=-1, keepdims=True)
    axis = divide_no_nan(rv, angle)
    s = np.sin(angle) * axis
    c = np.cos(angle)
    c1 = (1 - c) * axis
    ax, ay, az = axis[..., 0], axis[..., 1], axis[..., 2]
    m = np.empty(rv.shape[:-1] + (3, 3), dtype=F32)
    t = c1[..., 0] * ay
    m[..., 0, 1] = t - s[..., 2]
    m[..., 1, 0] = t + s[..., 2]
    t = c1[..., 0] * az
    m[..., 0, 2] = t + s[..., 1]
    m[..., 2, 0] = t - s[..., 1]
    t = c1[..., 1] * az
    m[..., 1, 2] = t - s[..., 0]
    m[..., 2, 1] = t + s[..., 0]
    d = c1 * axis + c
    m[..., 0, 0], m[..., 1, 1], m[..., 2, 2] = d[..., 0], d[..., 1], d[..., 2]
    return m


def mat2rotvec(R):
    """Quaternion-branch log map, pt/rotation.py:261-289 (branch order :280-285)."""
    R = np.asarray(R, dtype=F32)
    r00, r01, r02 = R[..., 0, 0], R[..., 0, 1], R[..., 0, 2]
    r10, r11, r12 = R[..., 1, 0], R[..., 1, 1], R[..., 1, 2]
    r20, r21, r22 = R[..., 2, 0], R[..., 2, 1], R[..., 2, 2]
    trace = r00 + r11 + r22
    one = F32(1.0)
    q0 = np.stack([r21 - r12, r02 - r20, r10 - r01, one + trace], -1)
    q1 = np.stack([(one - r22) + (r00 - r11), r10 + r01, r02 + r20, r21 - r12], -1)
    q2 = np.stack([r10 + r01, (one - r22) - (r00 - r11), r21 + r12, r02 - r20], -1)
    q3 = np.stack([r02 + r20, r21 + r12, (one + r22) - (r00 + r11), r10 - r01], -1)
    c0 = (trace > 0)[..., None]
    c1 = np.logical_and(r00 > r11, r00 > r22)[..., None]
    c2 = (r11 > r22)[..., None]
    q = np.where(c0, q0, np.where(c1, q1, np.where(c2, q2, q3)))
    xyz, w = q[..., :3], q[..., 3:]
    n = np.linalg.norm(xyz, axis=-1, keepdims=True)
    return (divide_no_nan(np.full_like(n, 2.0), n) * np.arctan2(n, w) * xyz).astype(F32)


def proj_so3(A):
    """Closest rotation by SVD with last-singular-vector flip, pt/rotation.py:100-110."""
    A = np.asarray(A, dtype=F32)
    U, _, Vh = np.linalg.svd(A)
    T = U @ Vh
    refl = (np.linalg.det(T) < 0)[..., None, None]
    Tm = T - 2 * U[..., :, -1:] @ Vh[..., -1:, :]
    return np.where(refl, Tm, T).astype(F32)


def align_unit_vectors(a, b):
    """Rotation taking unit a to unit b, pt/rotation.py:210-224."""
    cr = np.cross(a, b)
    dot = (a * b).sum(-1, keepdims=True)
    s = np.linalg.norm(cr, axis=-1, keepdims=True)
    ang = np.arctan2(s, dot)
    return rotvec2mat(divide_no_nan(cr * ang, s))


# --------------------------------------------------------------------------------------
# model + forward LBS (pt/bodymodel.py)
# --------------------------------------------------------------------------------------
class OracleModel:
    """float32 copy of the model constants (pt/bodymodel.py:80-93)."""

    def __init__(self, data, model_name='smpl'):
        self.model_name = model_name
        # always the float32-rounded constants (what the device holds), widened in exact mode
        f = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float32), dtype=F32)  # noqa: E731
        self.v_template = f(data.v_template)
        self.shapedirs = f(data.shapedirs)
        self.posedirs = f(data.posedirs)
        self.J_regressor_post_lbs = f(data.J_regressor_post_lbs)
        self.J_template = f(data.J_template)
        self.J_shapedirs = f(data.J_shapedirs)
        self.kid_shapedir = f(data.kid_shapedir)
        self.kid_J_shapedir = f(data.kid_J_shapedir)
        self.weights = f(data.weights)
        self.parents = [int(p) for p in data.kintree_parents]
        self.num_joints = int(data.num_joints)
        self.num_vertices = int(data.num_vertices)
        self.num_betas = self.shapedirs.shape[2]

    def forward(self, pose_rotvecs=None, shape_betas=None, trans=None, kid_factor=None,
                rel_rotmats=None, glob_rotmats=None, return_vertices=True):
        """pt/bodymodel.py:121-307."""
        n_rot = sum(x is not None for x in (pose_rotvecs, rel_rotmats, glob_rotmats))
        if n_rot > 1:
            raise ValueError('Only one rotation input may be provided')
        J, V = self.num_joints, self.num_vertices
        B = 0
        for arg in (pose_rotvecs, shape_betas, trans, rel_rotmats, glob_rotmats):
            if arg is not None:
                B = np.asarray(arg).shape[0]
                break
        if B == 0:
            out = dict(joints=np.zeros((0, J, 3), F32), orientations=np.zeros((0, J, 3, 3), F32))
            if return_vertices:
                out['vertices'] = np.zeros((0, V, 3), F32)
            return out
        par = self.parents
        if rel_rotmats is not None:
            rel = np.asarray(rel_rotmats, F32)
        elif pose_rotvecs is not None:
            rel = rotvec2mat(np.asarray(pose_rotvecs, F32).reshape(B, J, 3))
        elif glob_rotmats is None:
            rel = np.broadcast_to(np.eye(3, dtype=F32), (B, J, 3, 3)).copy()
        else:
            rel = None
        if glob_rotmats is None:
            glob = np.empty((B, J, 3, 3), F32)
            glob[:, 0] = rel[:, 0]
            for i in range(1, J):
                glob[:, i] = glob[:, par[i]] @ rel[:, i]
        else:
            glob = np.asarray(glob_rotmats, F32)
        par1 = np.array(par[1:])
        if rel is None:
            rel1 = np.swapaxes(glob[:, par1], -1, -2) @ glob[:, 1:]
        else:
            rel1 = rel[:, 1:]
        betas = np.zeros((B, 0), F32) if shape_betas is None else np.asarray(shape_betas, F32)
        nb = min(betas.shape[1], self.num_betas)
        kid = np.zeros((1,), F32) if kid_factor is None else np.asarray(kid_factor, F32).reshape(-1)
        j = (self.J_template[None]
             + np.einsum('jcs,bs->bjc', self.J_shapedirs[:, :, :nb], betas[:, :nb])
             + self.kid_J_shapedir[None] * kid[:, None, None])
        j = np.broadcast_to(j, (B, J, 3)).astype(F32)
        pos = np.empty((B, J, 3), F32)
        pos[:, 0] = j[:, 0]
        for i in range(1, J):
            bone = j[:, i] - j[:, par[i]]
            pos[:, i] = pos[:, par[i]] + np.einsum('bCc,bc->bC', glob[:, par[i]], bone)
        tr = np.zeros((1, 3), F32) if trans is None else np.asarray(trans, F32)
        if not return_vertices:
            return dict(joints=pos + tr[:, None], orientations=glob)
        feat = rel1.reshape(B, (J - 1) * 9)
        v_posed = (self.v_template[None]
                   + np.einsum('vcp,bp->bvc', self.shapedirs[:, :, :nb], betas[:, :nb])
                   + np.einsum('vcp,bp->bvc', self.posedirs, feat)
                   + self.kid_shapedir[None] * kid[:, None, None])
        transl = pos - np.einsum('bjCc,bjc->bjC', glob, j)
        rot_blend = np.einsum('vj,bjCc->bvCc', self.weights, glob)
        verts = np.einsum('bvCc,bvc->bvC', rot_blend, v_posed) + np.einsum('vj,bjC->bvC', self.weights, transl)
        return dict(joints=(pos + tr[:, None]).astype(F32), vertices=(verts + tr[:, None]).astype(F32),
                    orientations=glob)


# --------------------------------------------------------------------------------------
# fitter (pt/bodyfitter.py)
# --------------------------------------------------------------------------------------
class OraclePlan:
    """Independent re-derivation of the static masks (pt/bodyfitter.py:36-233)."""

    def __init__(self, model: OracleModel):
        J, V = model.num_joints, model.num_vertices
        par = model.parents
        self.smpl_family = model.model_name.startswith('smpl')
        part = np.argmax(model.weights, axis=1)
        if self.smpl_family:
            part = np.where(part == 10, 7, part)
            part = np.where(part == 11, 8, part)
        self.part = part.astype(np.int64)
        self.sel = [np.where(self.part == i)[0] for i in range(J)]
        self.cas = [[i] + [c for c in range(1, J) if par[c] == i] for i in range(J)]
        self.multi, self.bone, self.leaf = [], [], []
        for i in range(J):
            if self.smpl_family and i in (10, 11):
                continue
            n = len(self.cas[i])
            if n >= 3:
                self.multi.append(i)
            elif n == 2:
                self.bone.append(i)
            else:
                self.leaf.append(i)
        self.adjustable = [1, 2, 4, 5, 7, 8, 16, 17, 18, 19] if self.smpl_family else list(range(J))
        self.stat_parts = sorted(set(self.bone + self.leaf + self.adjustable))
        used = np.zeros(V, bool)
        for i in self.stat_parts:
            used[self.sel[i]] = True
        self.used = np.where(used)[0]


class OracleFitter:
    def __init__(self, model: OracleModel, enable_kid: bool = False):
        self.m = model
        self.enable_kid = enable_kid
        self.plan = OraclePlan(model)
        self.S = model.num_betas
        J, V = model.num_joints, model.num_vertices
        self.gram_supported = (not enable_kid) and J * 3 * V * self.S <= 2 ** 26
        self.template_mesh = model.forward(shape_betas=np.zeros((1, 0), F32))['vertices'][0]
        ext = [model.J_template[:, :, None], model.J_shapedirs]
        if enable_kid:
            ext.append(model.kid_J_shapedir[:, :, None])
        self.Jt_ext = np.concatenate(ext, axis=2).astype(F32)  # (J, 3, 1+S(+1))

    # ---- per-part statistics: pt/bodyfitter.py:235-280 (weights on the reference side) ----
    def part_sums(self, t, a, vw):
        B = max(t.shape[0], a.shape[0])
        J = self.m.num_joints
        raw = np.zeros((B, J, 3, 3), F32)
        s_t = np.zeros((B, J, 3), F32)
        s_a = np.zeros((B if vw is not None else a.shape[0], J, 3), F32)
        s_w = np.zeros((B if vw is not None else 1, J, 1), F32)
        for i in self.plan.stat_parts:
            sel = self.plan.sel[i]
            ti, ai = t[:, sel], a[:, sel]
            if vw is not None:
                wi = vw[:, sel, None]
                ai = ai * wi
                tsum = ti * wi
                s_w[:, i, 0] = vw[:, sel].sum(1)
            else:
                tsum = ti
                s_w[:, i, 0] = len(sel)
            raw[:, i] = np.einsum('bvi,bvj->bij', ti, np.broadcast_to(ai, (B,) + ai.shape[1:]))
            s_t[:, i] = tsum.sum(1)
            s_a[:, i] = ai.sum(1)
        return raw, s_t, s_a, s_w

    # ---- pt/bodyfitter.py:1321-1416 ----
    def fit_global_rotations(self, t, tj, a, aj, vw, jw):
        m, p = self.m, self.plan
        if tj is None or aj is None:
            tj = np.einsum('jv,bvc->bjc', m.J_regressor_post_lbs, t)
            aj = np.einsum('jv,bvc->bjc', m.J_regressor_post_lbs, a)
        B, J = t.shape[0], m.num_joints
        raw, s_t, s_a, s_w = self.part_sums(t, a, vw)
        R = np.zeros((B, J, 3, 3), F32)
        outer = lambda x, y: x[..., :, None] * y[..., None, :]  # noqa: E731
        for i in range(J):
            if p.smpl_family and i in (10, 11):
                continue
            cas = p.cas[i]
            mt = tj[:, cas].mean(1) if False else (tj[:, cas].sum(1) * F32(1.0 / len(cas)))
            ma = aj[:, cas].sum(1) * F32(1.0 / len(cas))
            if i in p.multi:
                rj = aj[:, cas]
                tsum = tj[:, cas]
                if jw is not None:
                    rj = rj * jw[:, cas, None]
                    tsum = tsum * jw[:, cas, None]
                    swj = jw[:, cas].sum(1)[:, None, None]
                else:
                    swj = F32(len(cas))
                rawj = np.einsum('bki,bkj->bij', tj[:, cas], np.broadcast_to(rj, (B,) + rj.shape[1:]))
                A = rawj - outer(tsum.sum(1), ma) - outer(mt, rj.sum(1)) + swj * outer(mt, ma)
                R[:, i] = proj_so3(A)
                continue
            A = raw[:, i] - outer(s_t[:, i], ma) - outer(mt, s_a[:, i]) + s_w[:, i, :, None] * outer(mt, ma)
            A = np.broadcast_to(A, (B, 3, 3))
            if i in p.leaf:
                R[:, i] = proj_so3(A)
                continue
            # bone part: swing from the bone direction, twist from the vertex covariance
            j0, j1 = cas[0], cas[1]
            b_ref = aj[:, j1] - aj[:, j0]
            b_tgt = tj[:, j1] - tj[:, j0]
            b_ref = divide_no_nan(b_ref, np.linalg.norm(b_ref, axis=-1, keepdims=True))
            b_tgt = divide_no_nan(b_tgt, np.linalg.norm(b_tgt, axis=-1, keepdims=True))
            Rs = align_unit_vectors(np.broadcast_to(b_ref, b_tgt.shape), b_tgt)
            H = Rs @ np.swapaxes(A, -1, -2)
            trH = H[:, 0, 0] + H[:, 1, 1] + H[:, 2, 2]
            bHb = np.einsum('bi,bij,bj->b', b_tgt, H, b_tgt)
            vee = np.stack([H[:, 1, 2] - H[:, 2, 1], H[:, 2, 0] - H[:, 0, 2], H[:, 0, 1] - H[:, 1, 0]], -1)
            ang = np.arctan2((b_tgt * vee).sum(-1), trH - bHb)
            R[:, i] = rotvec2mat(b_tgt * ang[:, None]) @ Rs
        if p.smpl_family:
            R[:, 10] = R[:, 7]
            R[:, 11] = R[:, 8]
        return R.astype(F32)

    # ---- shared front part of the shape solve: pt/bodyfitter.py:863-916 ----
    def _shape_front(self, glob):
        m = self.m
        B, J = glob.shape[0], m.num_joints
        par = m.parents
        rel = np.empty_like(glob)
        rel[:, 0] = glob[:, 0]
        for i in range(1, J):
            rel[:, i] = np.swapaxes(glob[:, par[i]], -1, -2) @ glob[:, i]
        n_ext = self.Jt_ext.shape[2]
        P = np.empty((B, J, 3, n_ext), F32)
        P[:, 0] = self.Jt_ext[0]
        for i in range(1, J):
            bone = self.Jt_ext[i] - self.Jt_ext[par[i]]
            P[:, i] = P[:, par[i]] + np.einsum('bCc,cs->bCs', glob[:, par[i]], bone)
        T = P - np.einsum('bjCc,jcs->bjCs', glob, self.Jt_ext)
        feat = rel[:, 1:].reshape(B, (J - 1) * 9)
        v_posed = m.v_template[None] + np.einsum('vcp,bp->bvc', m.posedirs, feat)
        return rel, P.astype(F32), T.astype(F32), v_posed.astype(F32)

    @staticmethod
    def _effective_weights(tj, vw, jw):
        """Shape-stage weight rule, pt/bodyfitter.py:1018-1028 / :1184-1189."""
        if tj is not None and vw is not None and jw is not None:
            return vw, jw
        if tj is None and vw is not None:
            return vw, None
        return None, None

    # ---- split-Gramian solve: pt/bodyfitter.py:960-1102 ----
    def fit_shape(self, glob, t, tj, vw, jw, reg, reg2, scale_reg=0.0, kid_reg=None,
                  scale_target=False, scale_fit=False, beta_ref=None, kid_ref=None, share_beta=False):
        if scale_target and scale_fit:
            raise ValueError('Only one of estim_scale_target and estim_scale_fit can be True')
        glob = np.asarray(glob, F32)
        rel, P, T, v_posed = self._shape_front(glob)
        if share_beta:
            return self._fit_shape_general(glob, rel, P, T, v_posed, t, tj, vw, jw, reg, reg2, scale_reg,
                                           kid_reg, scale_target, scale_fit, beta_ref, kid_ref, share_beta=True)
        if self.gram_supported and not (scale_target or scale_fit):
            return self._fit_shape_gram(glob, rel, P, T, v_posed, t, tj, vw, jw, reg, reg2, beta_ref)
        return self._fit_shape_general(glob, rel, P, T, v_posed, t, tj, vw, jw, reg, reg2, scale_reg,
                                       kid_reg, scale_target, scale_fit, beta_ref, kid_ref)

    def _fit_shape_gram(self, glob, rel, P, T, v_posed, t, tj, vw, jw, reg, reg2, beta_ref):
        m = self.m
        B, S = t.shape[0], self.S
        rot_blend = np.einsum('vj,bjCc->bvCc', m.weights, glob)
        pos = np.einsum('bvCc,bvc->bvC', rot_blend, v_posed) + np.einsum('vj,bjC->bvC', m.weights, T[..., 0])
        jac = (np.einsum('bvCc,vcs->bvCs', rot_blend, m.shapedirs)
               + np.einsum('vj,bjCs->bvCs', m.weights, T[..., 1:]))
        b = t - pos
        evw, ejw = self._effective_weights(tj, vw, jw)
        D = np.float64

        def block(jac_, b_, w_):
            n = jac_.shape[1]
            if w_ is None:
                wj = jac_
                wsum = np.full((B,), float(n), D)
                sb = b_.sum(1)
            else:
                wj = jac_ * w_[:, :, None, None]
                wsum = w_.sum(1).astype(D)
                sb = (b_ * w_[:, :, None]).sum(1)
            G = np.einsum('bvcs,bvct->bst', wj, jac_)  # float32 products and sums
            r = np.einsum('bvcs,bvc->bs', wj, b_)
            sA = wj.sum(1)
            return G.astype(D), r.astype(D), sA.astype(D), sb.astype(D), wsum

        G, r, sA, sb, W = block(jac, b, evw)
        if tj is not None:
            Gj, rj, sAj, sbj, Wj = block(P[..., 1:], tj - P[..., 0], ejw)
            G, r, sA, sb, W = G + Gj, r + rj, sA + sAj, sb + sbj, W + Wj
        Ws = np.where(W == 0, 1.0, W)[:, None, None]
        Gc = G - np.einsum('bcs,bct->bst', sA, sA) / Ws
        rc = r - np.einsum('bcs,bc->bs', sA, sb) / Ws[:, :, 0]
        lam = np.concatenate([np.full(2, float(reg2)), np.full(S - 2, float(reg))]).astype(D)
        ref = np.zeros((B, S), D)
        if beta_ref is not None:
            br = np.asarray(beta_ref, D)
            n = min(br.shape[1], S)
            ref[:, :n] = br[:, :n]
        x = np.linalg.solve(Gc + np.diag(lam)[None], (rc + lam * ref)[..., None])[..., 0]
        mean_A = sA / Ws
        mean_b = sb / Ws[:, :, 0]
        trans = (mean_b - np.einsum('bcs,bs->bc', mean_A, x)).astype(F32)
        beta = x.astype(F32)
        joints = P[..., 0] + np.einsum('bjcs,bs->bjc', P[..., 1:], beta) + trans[:, None]
        verts = pos + np.einsum('bvcs,bs->bvc', jac, beta) + trans[:, None]
        return dict(shape_betas=beta, trans=trans, relative_orientations=rel,
                    joints=joints.astype(F32), vertices=verts.astype(F32))

    # ---- general solve (kid / scale unknowns, float32): pt/bodyfitter.py:1104-1319, pt/lstsq.py:7-29 ----
    def _fit_shape_general(self, glob, rel, P, T, v_posed, t, tj, vw, jw, reg, reg2, scale_reg,
                           kid_reg, scale_target, scale_fit, beta_ref, kid_ref, share_beta=False):
        m = self.m
        B, S = t.shape[0], self.S
        rot_blend = np.einsum('vj,bjCc->bvCc', m.weights, glob)
        sd = m.shapedirs
        if self.enable_kid:
            sd = np.concatenate([sd, m.kid_shapedir[:, :, None]], axis=2)
        ext = np.concatenate([np.einsum('bvCc,bvc->bvC', rot_blend, v_posed)[..., None],
                              np.einsum('bvCc,vcs->bvCs', rot_blend, sd)], axis=3)
        ext = ext + np.einsum('vj,bjCs->bvCs', m.weights, T)
        if tj is None:
            tgt, pos, jac = t, ext[..., 0], ext[..., 1:]
        else:
            tgt = np.concatenate([t, tj], 1)
            pos = np.concatenate([ext[..., 0], P[..., 0]], 1)
            jac = np.concatenate([ext[..., 1:], P[..., 1:]], 1)
        if scale_target:
            A = np.concatenate([jac, -tgt[..., None]], 3)
        elif scale_fit:
            A = np.concatenate([jac, pos[..., None]], 3)
        else:
            A = jac
        b = tgt - pos
        evw, ejw = self._effective_weights(tj, vw, jw)
        if evw is None:
            w = np.ones(A.shape[:2], F32)
        elif tj is not None:
            w = np.concatenate([evw, ejw], 1).astype(F32)
        else:
            w = evw.astype(F32)
        wsum = w.sum(1)[:, None, None]
        safe = np.where(wsum == 0, 1, wsum)
        mean_A = np.where(wsum[..., None] == 0, 0, (w[:, :, None, None] * A).sum(1, keepdims=True) / safe[..., None])
        mean_b = np.where(wsum == 0, 0, (w[:, :, None] * b).sum(1, keepdims=True) / safe)
        A = (A - mean_A).astype(F32)
        b = (b - mean_b).astype(F32)
        npar = A.shape[-1]
        A2 = A.reshape(B, -1, npar)
        b2 = b.reshape(B, -1)
        w3 = np.repeat(w, 3, axis=1)
        lam = [float(reg2)] * 2 + [float(reg)] * (S - 2)
        ref = np.zeros((B, S), F32)
        if beta_ref is not None:
            br = np.asarray(beta_ref, F32)
            n = min(br.shape[1], S)
            ref[:, :n] = br[:, :n]
        if self.enable_kid:
            lam.append(float(reg if kid_reg is None else kid_reg))
            kr = np.zeros((B,), F32) if kid_ref is None else np.asarray(kid_ref, F32)
            ref = np.concatenate([ref, kr[:, None]], 1)
        if scale_target or scale_fit:
            lam.append(float(scale_reg))
            ref = np.concatenate([ref, np.zeros((B, 1), F32)], 1)
        lam = np.array(lam, F32)
        WA = w3[:, :, None] * A2
        G = np.einsum('bns,bnt->bst', WA, A2) + np.diag(lam)[None]
        rhs = np.einsum('bns,bn->bs', WA, b2) + lam * ref
        n_sh = S + (1 if self.enable_kid else 0)
        if share_beta and npar > n_sh:
            # pt/lstsq.py:32-90 lstsq_partial_share: betas (+ kid) shared over the batch, the scale column per instance.
            # The regulariser enters as extra rows (row e_i, weight lambda_i, right-hand side lambda_i ref_i), so its
            # reference term is lambda_i^2 ref_i here; the independent column is eliminated per instance (Schur
            # complement), the shared system is summed over the batch, then the independent unknown is recovered.
            rp = np.einsum('bns,bn->bs', WA, b2) + (lam ** 2) * ref
            Gss, Gsz, Gzz = G[:, :n_sh, :n_sh], G[:, :n_sh, n_sh:], G[:, n_sh:, n_sh:]
            c_s = np.linalg.solve(Gzz, np.swapaxes(Gsz, 1, 2))            # (B, n_indep, n_sh)
            c_r = np.linalg.solve(Gzz, rp[:, n_sh:, None])                  # (B, n_indep, 1)
            Sb_ = (Gss - Gsz @ c_s).sum(0)
            tb_ = (rp[:, :n_sh, None] - Gsz @ c_r).sum(0)
            x_s = np.linalg.solve(Sb_.astype(F32), tb_.astype(F32))       # (n_sh, 1)
            x_z = c_r - c_s @ x_s[None]
            x = np.concatenate([np.broadcast_to(x_s[None, :, 0], (B, n_sh)), x_z[:, :, 0]], 1).astype(F32)
        elif share_beta:
            # pt/lstsq.py:43-45 -> lstsq(..., shared=True): diag(lambda) is inside the per-instance
            # Gramian that gets summed over the batch, and the regulariser-reference term is not passed
            Gs = G.sum(0)
            rs_ = np.einsum('bns,bn->bs', WA, b2).sum(0)
            x = np.broadcast_to(np.linalg.solve(Gs.astype(F32), rs_.astype(F32)), (B, npar)).astype(F32)
        else:
            x = np.linalg.solve(G.astype(F32), rhs.astype(F32)[..., None])[..., 0].astype(F32)
        trans = (mean_b[:, 0] - np.einsum('bcs,bs->bc', mean_A[:, 0], x)).astype(F32)
        beta = x[:, :S]
        out = dict(shape_betas=beta, trans=trans, relative_orientations=rel)
        kid = None
        if self.enable_kid:
            kid = x[:, S]
            out['kid_factor'] = kid
        if scale_target or scale_fit:
            sc = x[:, -1] + 1
            if scale_fit:
                beta = beta / sc[:, None]
                if kid is not None:
                    kid = kid / sc
            out['scale_corr'] = sc.astype(F32)
        full = beta if kid is None else np.concatenate([beta, kid[:, None]], 1)
        out['joints'] = (P[..., 0] + np.einsum('bjcs,bs->bjc', P[..., 1:], full) + trans[:, None]).astype(F32)
        out['vertices'] = (ext[..., 0] + np.einsum('bvcs,bs->bvc', ext[..., 1:], full) + trans[:, None]).astype(F32)
        return out

    # ---- final adjustment, sequential form: pt/bodyfitter.py:1418-1469, :1546-1595 ----
    def fit_global_rotations_dependent(self, t, tj, a, aj, vw, jw, R_prev, betas, scale_corr, trans, kid):
        m, p = self.m, self.plan
        true_aj = aj
        if tj is None or aj is None:
            tj = np.einsum('jv,bvc->bjc', m.J_regressor_post_lbs, t)
            aj = np.einsum('jv,bvc->bjc', m.J_regressor_post_lbs, a)
        if true_aj is None:
            true_aj = aj
        B, J = t.shape[0], m.num_joints
        par = m.parents
        j = m.J_template[None] + np.einsum('jcs,bs->bjc', m.J_shapedirs, betas[:, : self.S])
        if kid is not None:
            j = j + m.kid_J_shapedir[None] * np.asarray(kid, F32)[:, None, None]
        if scale_corr is not None:
            j = j * scale_corr
        j = j.astype(F32)
        raw, s_t, s_a, s_w = self.part_sums(t, a, vw)
        outer = lambda x, y: x[..., :, None] * y[..., None, :]  # noqa: E731
        R = np.array(R_prev, F32)
        pos = np.empty((B, J, 3), F32)
        for i in range(J):
            if i == 0:
                pos[:, 0] = j[:, 0] + trans
            else:
                pos[:, i] = pos[:, par[i]] + np.einsum('bCc,bc->bC', R[:, par[i]], j[:, i] - j[:, par[i]])
            if p.smpl_family and i in (10, 11):
                R[:, i] = R[:, 7 if i == 10 else 8]
                continue
            if i not in p.adjustable:
                continue
            c_t, c_a = pos[:, i], true_aj[:, i]
            A = raw[:, i] - outer(s_t[:, i], c_a) - outer(c_t, s_a[:, i]) + s_w[:, i, :, None] * outer(c_t, c_a)
            cas = p.cas[i]
            ej = tj[:, cas] - c_t[:, None]
            dj = aj[:, cas] - c_a[:, None]
            if jw is not None:
                dj = dj * jw[:, cas, None]
            A = A + np.einsum('bki,bkj->bij', ej, np.broadcast_to(dj, ej.shape))
            R[:, i] = proj_so3(A) @ R_prev[:, i]
        return R

    # ---- driver: pt/bodyfitter.py:283-549 ----
    def fit(self, target_vertices, target_joints=None, vertex_weights=None, joint_weights=None,
            num_iter=1, beta_regularizer=1, beta_regularizer2=0, scale_regularizer=0,
            kid_regularizer=None, share_beta=False, final_adjust_rots=True, scale_target=False,
            scale_fit=False, initial_pose_rotvecs=None, initial_shape_betas=None,
            initial_kid_factor=None, requested_keys=None):
        if requested_keys is None:
            requested_keys = ['pose_rotvecs']
        m = self.m
        t = np.asarray(target_vertices, F32)
        tj = None if target_joints is None else np.asarray(target_joints, F32)
        vw = None if vertex_weights is None else np.asarray(vertex_weights, F32)
        jw = None if joint_weights is None else np.asarray(joint_weights, F32)
        if tj is None:
            mean = t.mean(1)
            t = t - mean[:, None]
        else:
            mean = np.concatenate([t, tj], 1).mean(1)
            t = t - mean[:, None]
            tj = tj - mean[:, None]
        if initial_pose_rotvecs is not None or initial_shape_betas is not None:
            init = m.forward(pose_rotvecs=initial_pose_rotvecs, shape_betas=initial_shape_betas,
                             kid_factor=initial_kid_factor)
            glob = self.fit_global_rotations(t, tj, init['vertices'], init['joints'], vw, jw) @ init['orientations']
        else:
            glob = self.fit_global_rotations(t, tj, self.template_mesh[None], m.J_template[None], vw, jw)
        for _ in range(num_iter - 1):
            res = self.fit_shape(glob, t, tj, vw, jw, beta_regularizer, beta_regularizer2, 0.0,
                                 kid_regularizer, False, False, initial_shape_betas, initial_kid_factor, share_beta)
            rj = res['joints'] if tj is not None else None
            glob = self.fit_global_rotations(t, tj, res['vertices'], rj, vw, jw) @ glob
        res = self.fit_shape(glob, t, tj, vw, jw, beta_regularizer, beta_regularizer2, scale_regularizer,
                             kid_regularizer, scale_target, scale_fit, initial_shape_betas, initial_kid_factor, share_beta)
        ref_v = res['vertices']
        ref_j = res['joints'] if (tj is not None or final_adjust_rots) else None
        kid = res['kid_factor'] if self.enable_kid else None
        sc = res['scale_corr'][:, None, None] if (scale_target or scale_fit) else None
        if final_adjust_rots:
            if scale_target:
                glob = self.fit_global_rotations_dependent(
                    t * sc, tj * sc if tj is not None else None, ref_v, ref_j, vw, jw, glob,
                    res['shape_betas'], None, res['trans'], kid)
            elif scale_fit:
                tr = res['trans'][:, None]
                glob = self.fit_global_rotations_dependent(
                    t, tj, sc * ref_v + (1 - sc) * tr, sc * ref_j + (1 - sc) * tr, vw, jw, glob,
                    res['shape_betas'], sc, res['trans'], kid)
            else:
                glob = self.fit_global_rotations_dependent(
                    t, tj, ref_v, ref_j, vw, jw, glob, res['shape_betas'], None, res['trans'], kid)
        out = dict(shape_betas=res['shape_betas'])
        if scale_target:
            out['trans'] = res['trans'] + mean * res['scale_corr'][:, None]
        elif scale_fit:
            out['trans'] = res['trans'] + mean / res['scale_corr'][:, None]
        else:
            out['trans'] = res['trans'] + mean
        out['orientations'] = glob
        out['relative_orientations'] = res['relative_orientations']
        if 'relative_orientations' in requested_keys or 'pose_rotvecs' in requested_keys:
            par = m.parents
            rel = np.empty_like(glob)
            rel[:, 0] = glob[:, 0]
            for i in range(1, m.num_joints):
                rel[:, i] = np.swapaxes(glob[:, par[i]], -1, -2) @ glob[:, i]
            out['relative_orientations'] = rel
        if 'pose_rotvecs' in requested_keys:
            out['pose_rotvecs'] = mat2rotvec(out['relative_orientations']).reshape(glob.shape[0], -1)
        if kid is not None:
            out['kid_factor'] = kid
        if sc is not None:
            out['scale_corr'] = res['scale_corr']
        return out

    # ---- pt/bodyfitter.py:552-653 ----
    def fit_with_known_pose(self, pose_rotvecs, target_vertices, target_joints=None, vertex_weights=None,
                            joint_weights=None, beta_regularizer=1, beta_regularizer2=0,
                            scale_regularizer=0, kid_regularizer=None, scale_target=False,
                            scale_fit=False, beta_regularizer_reference=None,
                            kid_regularizer_reference=None, share_beta=False):
        m = self.m
        t = np.asarray(target_vertices, F32)
        tj = None if target_joints is None else np.asarray(target_joints, F32)
        if tj is None:
            mean = t.mean(1)
            t = t - mean[:, None]
        else:
            mean = np.concatenate([t, tj], 1).mean(1)
            t, tj = t - mean[:, None], tj - mean[:, None]
        glob = m.forward(pose_rotvecs=pose_rotvecs, return_vertices=False)['orientations']
        res = self.fit_shape(glob, t, tj, vertex_weights, joint_weights, beta_regularizer,
                             beta_regularizer2, scale_regularizer, kid_regularizer, scale_target,
                             scale_fit, beta_regularizer_reference, kid_regularizer_reference, share_beta)
        res['trans'] = res['trans'] + mean
        res.pop('vertices')
        res.pop('joints')
        return res


def fit_scale_and_translation(t, a, tj, aj, vw=None, jw=None, scale=False):
    """pt/bodyfitter.py:1628-1681."""
    if tj is None or aj is None:
        tb, ab = t, a
        w = vw if vw is not None else np.ones(t.shape[:2], F32)
    else:
        tb, ab = np.concatenate([t, tj], 1), np.concatenate([a, aj], 1)
        w = np.concatenate([vw, jw], 1) if (vw is not None and jw is not None) else np.ones(tb.shape[:2], F32)
    w = w / w.sum(1, keepdims=True)
    mt = (tb * w[..., None]).sum(1)
    ma = (ab * w[..., None]).sum(1)
    if scale:
        sst = (((tb - mt[:, None]) ** 2) * w[..., None]).sum((1, 2))
        ssa = (((ab - ma[:, None]) ** 2) * w[..., None]).sum((1, 2))
        sc = np.sqrt(sst / ssa)
        return sc, mt - sc[:, None] * ma
    return None, mt - ma


def fit_with_known_shape(fitter, shape_betas, target_vertices, target_joints=None, vertex_weights=None,
                         joint_weights=None, kid_factor=None, num_iter=1, final_adjust_rots=True,
                         initial_pose_rotvecs=None, scale_fit=False, requested_keys=None):
    """pt/bodyfitter.py:656-838 (``fitter`` is an OracleFitter)."""
    if requested_keys is None:
        requested_keys = ['pose_rotvecs']
    m = fitter.m
    t = np.asarray(target_vertices, F32)
    tj = None if target_joints is None else np.asarray(target_joints, F32)
    vw, jw = vertex_weights, joint_weights
    if tj is None:
        mean = t.mean(1)
        t = t - mean[:, None]
    else:
        mean = np.concatenate([t, tj], 1).mean(1)
        t, tj = t - mean[:, None], tj - mean[:, None]
    B = t.shape[0]
    init = m.forward(shape_betas=shape_betas, kid_factor=kid_factor, pose_rotvecs=initial_pose_rotvecs)
    bc = lambda x: np.broadcast_to(x, (B,) + x.shape[1:])  # noqa: E731
    glob = fitter.fit_global_rotations(t, tj, bc(init['vertices']), bc(init['joints']), vw, jw) @ bc(init['orientations'])
    for _ in range(num_iter - 1):
        res = m.forward(glob_rotmats=glob, shape_betas=shape_betas, kid_factor=kid_factor)
        glob = fitter.fit_global_rotations(t, tj, res['vertices'], res['joints'] if tj is not None else None, vw, jw) @ glob
    res = m.forward(glob_rotmats=glob, shape_betas=shape_betas, kid_factor=kid_factor)
    rv, rj = res['vertices'], res['joints']
    sc, tr = fit_scale_and_translation(t, rv, tj, rj, vw, jw, scale=scale_fit)
    if final_adjust_rots:
        if scale_fit:
            s3 = sc[:, None, None]
            glob = fitter.fit_global_rotations_dependent(t, tj, s3 * rv + tr[:, None], s3 * rj + tr[:, None], vw, jw,
                                                         glob, shape_betas, s3, tr, kid_factor)
        else:
            glob = fitter.fit_global_rotations_dependent(t, tj, rv + tr[:, None], rj + tr[:, None], vw, jw, glob,
                                                         shape_betas, None, tr, kid_factor)
    out = dict(trans=tr + mean, orientations=glob)
    if scale_fit:
        out['scale_corr'] = sc
    if 'relative_orientations' in requested_keys or 'pose_rotvecs' in requested_keys:
        par = m.parents
        rel = np.empty_like(glob)
        rel[:, 0] = glob[:, 0]
        for i in range(1, m.num_joints):
            rel[:, i] = np.swapaxes(glob[:, par[i]], -1, -2) @ glob[:, i]
        out['relative_orientations'] = rel
        if 'pose_rotvecs' in requested_keys:
            out['pose_rotvecs'] = mat2rotvec(rel).reshape(B, -1)
    return out


def convert_vertices_csr(indptr, indices, data, verts):
    """CSR (V_out x V_in) applied to (B, V_in, 3): pt/bodyconverter.py:129-149."""
    verts = np.asarray(verts, F32)
    V_out = len(indptr) - 1
    out = np.zeros((verts.shape[0], V_out, 3), F32)
    for r in range(V_out):
        for k in range(indptr[r], indptr[r + 1]):
            out[:, r] += F32(data[k]) * verts[:, indices[k]]
    return out
