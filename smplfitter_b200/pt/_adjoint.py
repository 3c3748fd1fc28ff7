"""Gradient path of ``BodyModel.forward``, ``BodyFitter.fit`` / ``fit_with_known_pose`` / ``fit_with_known_shape``
(reference feature: README.md:13, tests/pt/test_fitter_grad.py:31-99 -- backprop through the fit must give finite
gradients that agree with finite differences).

The VALUES always come from the CUDA kernels.  Only when an input of the ``smplfit_b200::forward`` / ``::fit`` custom
op (or of a known-pose / known-shape call, wrapped by ``differentiable_call``) requires grad does autograd call the
backward registered here, which re-evaluates the same closed-form algorithm with differentiable torch operations on
slices of the batch (instances are independent; ``share_beta`` couples them: one slice) and pulls the incoming output
gradients back through that evaluation.  It is a recompute-in-backward adjoint: nothing here runs on the inference
path (``tests/test_host_cpu.py::test_gradient_evaluation_is_reachable_only_from_backward_paths``,
``tests/test_gpu_grad.py::test_inference_never_evaluates_the_adjoint``), nothing is kept alive between forward and
backward except the inputs, and the memory of the intermediate Jacobians is bounded by the slice size instead of
the batch size.

Two pieces are hand-derived instead of left to torch's autograd because the stock derivatives are singular exactly
where the fit operates:

* ``proj_so3`` (closest rotation, pt/rotation.py:100-110): torch's SVD backward divides by differences of singular
  values (isotropic part covariances make them equal); the derivative of the PRODUCT U D V^T only involves sums
  s_i + s_j, see ``_ProjSO3.backward``.
* ``rotvec2mat`` / ``mat2rotvec`` use series forms near the identity so that the derivative exists at zero rotation.

Shaped by a profiler pass on the B200 (``scripts/grad_profile.py``): 3x3 products are broadcast multiply-sums
(``_mm3`` / ``_mv3``; batched BLAS ran them as one tiny GEMV per matrix), the joint chains are path sums over an
ancestor matrix or one product per tree level instead of a loop over the joints.

What is differentiable: ``forward`` with respect to every tensor input; the three fits with respect to the targets,
the weights, the initial guesses and the known pose / betas, for every option (joints or not, weights, ``num_iter``,
``final_adjust_rots``, the regularisers and their references, ``enable_kid``, ``scale_target`` / ``scale_fit``,
``share_beta``).  It is the same function as the CUDA path's: in float64 it reproduces the float64 evaluation stored
with the golden cases (their ``exact_*`` arrays) to 1e-13 (``tests/test_adjoint_cpu.py``).
"""

from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------------------------------
# rotations
# ----------------------------------------------------------------------------------------------------------------------
def _mm3(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """(..., 3, 3) @ (..., 3, n) as a broadcast multiply + sum: batched BLAS calls on 3x3 operands run as one tiny GEMV
    per matrix (measured on the B200: 95 % of the backward's GPU time before this), elementwise kernels do not."""
    return (A.unsqueeze(-1) * B.unsqueeze(-3)).sum(-2)


def _mv3(A: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """(..., 3, 3) @ (..., 3)."""
    return (A * v.unsqueeze(-2)).sum(-1)


def _skew(v: torch.Tensor) -> torch.Tensor:
    z = torch.zeros_like(v[..., 0])
    return torch.stack([z, -v[..., 2], v[..., 1], v[..., 2], z, -v[..., 0], -v[..., 1], v[..., 0], z], -1).reshape(
        v.shape[:-1] + (3, 3))


def rotvec2mat(rv: torch.Tensor) -> torch.Tensor:
    """Rodrigues (pt/rotation.py:236-258) as I + a K + b K^2 with a = sin(t)/t, b = (1-cos t)/t^2 and their Taylor
    forms below t^2 = 1e-8, so that value and derivative exist at t = 0."""
    t2 = (rv * rv).sum(-1)
    small = t2 < 1e-8
    t2s = torch.where(small, torch.ones_like(t2), t2)
    t = t2s.sqrt()
    a = torch.where(small, 1 - t2 / 6, torch.sin(t) / t)
    b = torch.where(small, 0.5 - t2 / 24, (1 - torch.cos(t)) / t2s)
    K = _skew(rv)
    eye = torch.eye(3, dtype=rv.dtype, device=rv.device)
    return eye + a[..., None, None] * K + b[..., None, None] * _mm3(K, K)


def mat2rotvec(R: torch.Tensor) -> torch.Tensor:
    """Log map through the largest-component quaternion (pt/rotation.py:261-289, branch order :280-285)."""
    r00, r01, r02 = R[..., 0, 0], R[..., 0, 1], R[..., 0, 2]
    r10, r11, r12 = R[..., 1, 0], R[..., 1, 1], R[..., 1, 2]
    r20, r21, r22 = R[..., 2, 0], R[..., 2, 1], R[..., 2, 2]
    tr = r00 + r11 + r22
    q0 = torch.stack([r21 - r12, r02 - r20, r10 - r01, 1 + tr], -1)
    q1 = torch.stack([(1 - r22) + (r00 - r11), r10 + r01, r02 + r20, r21 - r12], -1)
    q2 = torch.stack([r10 + r01, (1 - r22) - (r00 - r11), r21 + r12, r02 - r20], -1)
    q3 = torch.stack([r02 + r20, r21 + r12, (1 + r22) - (r00 + r11), r10 - r01], -1)
    c0 = (tr > 0)[..., None]
    c1 = ((r00 > r11) & (r00 > r22))[..., None]
    c2 = (r11 > r22)[..., None]
    q = torch.where(c0, q0, torch.where(c1, q1, torch.where(c2, q2, q3)))
    xyz, w = q[..., :3], q[..., 3]
    n2 = (xyz * xyz).sum(-1)
    small = n2 < 1e-16 * w * w  # angle -> 0: 2 atan2(n, w) / n -> 2 / w
    n = torch.where(small, torch.ones_like(n2), n2).sqrt()
    f = torch.where(small, 2 / torch.where(small, w, torch.ones_like(w)), 2 * torch.atan2(n, w) / n)
    return f[..., None] * xyz


class _ProjSO3(torch.autograd.Function):
    """R = U diag(1, 1, det(U V^T)) V^T for A = U S V^T.

    With the signed factorisation A = U S' V'^T (V' = V D, S' = S D) R = U V'^T is the orthogonal polar factor of A, so
    dR = U X V'^T with X antisymmetric and X_ij (s'_i + s'_j) = M_ij - M_ji, M = U^T dA V'.  The pull-back of a
    cotangent G is therefore U Y V'^T with Y_ij = (N_ij - N_ji) / (s'_i + s'_j), N = U^T G V'."""

    @staticmethod
    def forward(ctx, A):
        U, S, Vh = torch.linalg.svd(A)
        d = torch.sign(torch.linalg.det(_mm3(U, Vh)))
        d = torch.where(d == 0, torch.ones_like(d), d)
        D = torch.ones_like(S)
        D[..., 2] = d
        Vp = Vh.transpose(-1, -2) * D[..., None, :]
        ctx.save_for_backward(U, S * D, Vp)
        return _mm3(U, Vp.transpose(-1, -2))

    @staticmethod
    def backward(ctx, G):
        U, Sp, Vp = ctx.saved_tensors
        N = _mm3(_mm3(U.transpose(-1, -2), G), Vp)
        c = Sp[..., :, None] + Sp[..., None, :]
        tiny = torch.finfo(G.dtype).eps * Sp[..., :1, None].abs().clamp_min(torch.finfo(G.dtype).tiny)
        c = torch.where(c.abs() < tiny, torch.where(c < 0, -tiny, tiny), c)
        Y = (N - N.transpose(-1, -2)) / c
        return _mm3(_mm3(U, Y), Vp.transpose(-1, -2))


def proj_so3(A: torch.Tensor) -> torch.Tensor:
    return _ProjSO3.apply(A)


def _unit(v: torch.Tensor) -> torch.Tensor:
    n2 = (v * v).sum(-1, keepdim=True)
    ok = n2 > 0
    return torch.where(ok, v / torch.where(ok, n2, torch.ones_like(n2)).sqrt(), torch.zeros_like(v))


def _align_unit_vectors(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Rotation taking unit a to unit b about a x b (pt/rotation.py:210-224)."""
    cr = torch.linalg.cross(a, b)
    dot = (a * b).sum(-1, keepdim=True)
    s2 = (cr * cr).sum(-1, keepdim=True)
    ok = s2 > 1e-30
    s = torch.where(ok, s2, torch.ones_like(s2)).sqrt()
    rv = torch.where(ok, cr * (torch.atan2(s, dot) / s), cr)
    return rotvec2mat(rv)


def _outer(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return x[..., :, None] * y[..., None, :]


# ----------------------------------------------------------------------------------------------------------------------
# constants of one model in the dtype / on the device of the gradient evaluation
# ----------------------------------------------------------------------------------------------------------------------
class Constants:
    def __init__(self, bm, dtype: torch.dtype, device: torch.device):
        f = lambda x: x.detach().to(device=device, dtype=dtype)  # noqa: E731
        self.v_template, self.shapedirs, self.posedirs = f(bm.v_template), f(bm.shapedirs), f(bm.posedirs)
        self.J_regressor = f(bm.J_regressor_post_lbs)
        self.J_template, self.J_shapedirs = f(bm.J_template), f(bm.J_shapedirs)
        self.kid_shapedir, self.kid_J_shapedir = f(bm.kid_shapedir), f(bm.kid_J_shapedir)
        self.weights = f(bm.weights)
        self.parents = [int(p) for p in bm.kintree_parents]
        self.J, self.V, self.S = bm.num_joints, bm.num_vertices, bm.num_betas
        p = bm._plan
        self.plan = p
        idx = lambda a: torch.as_tensor(a, dtype=torch.int64, device=device)  # noqa: E731
        self.used = idx(p.used_vertex_indices)
        self.part_of_used = idx(p.part_assignment[p.used_vertex_indices])
        self.part_counts = torch.as_tensor(p.part_counts, dtype=dtype, device=device)
        self.center = torch.as_tensor(p.center_matrix, dtype=dtype, device=device)
        self.mjp = torch.as_tensor(p.mjp_joint_membership, dtype=dtype, device=device)
        self.multi, self.bone, self.leaf = idx(p.multi_joint_parts), idx(p.bone_parts), idx(p.leaf_parts)
        self.bone_pairs = idx(p.bone_pairs)
        self.assemble = idx(p.assemble_indices)
        self.par1 = idx(self.parents[1:])
        depth = [0] * self.J
        for i in range(1, self.J):
            depth[i] = depth[self.parents[i]] + 1
        self.levels = [[i for i in range(self.J) if depth[i] == d] for d in range(1, max(depth) + 1)]
        anc = np.zeros((self.J, self.J), np.float64)  # anc[i, k]: joint k >= 1 lies on the path root -> i
        for i in range(1, self.J):
            anc[i] = anc[self.parents[i]]
            anc[i, i] = 1.0
        self.anc1 = torch.as_tensor(anc[:, 1:], dtype=dtype, device=device)
        self.template_mesh = None  # filled on first use (forward at zero pose / shape)


def constants(bm, dtype: torch.dtype, device: torch.device) -> Constants:
    cache = bm.__dict__.setdefault('_adjoint_constants', {})
    key = (dtype, str(device))
    if key not in cache:
        cache[key] = Constants(bm, dtype, device)
    return cache[key]


# ----------------------------------------------------------------------------------------------------------------------
# forward LBS (pt/bodymodel.py:121-307)
# ----------------------------------------------------------------------------------------------------------------------
def _chain(c: Constants, rel: torch.Tensor) -> torch.Tensor:
    """Global orientations from relative ones, one batched product per level of the kinematic tree."""
    glob: List[Optional[torch.Tensor]] = [None] * c.J
    glob[0] = rel[:, 0]
    for level in c.levels:
        parent = torch.stack([glob[c.parents[i]] for i in level], 1)
        new = _mm3(parent, rel[:, level])
        for n, i in enumerate(level):
            glob[i] = new[:, n]
    return torch.stack(glob, 1)


def _relative(c: Constants, glob: torch.Tensor) -> torch.Tensor:
    return torch.cat([glob[:, :1], _mm3(glob[:, c.par1].transpose(-1, -2), glob[:, 1:])], 1)


def _positions(c: Constants, glob: torch.Tensor, j: torch.Tensor) -> torch.Tensor:
    """Joint positions of the chain: root + sum over the path of the parent-rotated bones (no loop over joints)."""
    bones = _mv3(glob[:, c.par1], j[:, 1:] - j[:, c.par1])
    return j[:, :1] + torch.einsum('ik,bkc->bic', c.anc1, bones)


def lbs(c: Constants, pose_rotvecs=None, shape_betas=None, trans=None, kid_factor=None, rel_rotmats=None,
        glob_rotmats=None, return_vertices: bool = True):
    B = 1
    for x in (pose_rotvecs, shape_betas, trans, rel_rotmats, glob_rotmats):
        if x is not None:
            B = x.shape[0]
            break
    dt, dev = c.v_template.dtype, c.v_template.device
    if rel_rotmats is not None:
        rel = rel_rotmats
    elif pose_rotvecs is not None:
        rel = rotvec2mat(pose_rotvecs.reshape(B, c.J, 3))
    elif glob_rotmats is None:
        rel = torch.eye(3, dtype=dt, device=dev).expand(B, c.J, 3, 3)
    else:
        rel = None
    glob = _chain(c, rel) if glob_rotmats is None else glob_rotmats
    rel1 = _relative(c, glob)[:, 1:] if rel is None else rel[:, 1:]
    nb = 0 if shape_betas is None else min(shape_betas.shape[1], c.S)
    j = c.J_template[None]
    if nb:
        j = j + torch.einsum('jcs,bs->bjc', c.J_shapedirs[:, :, :nb], shape_betas[:, :nb])
    if kid_factor is not None:
        j = j + c.kid_J_shapedir[None] * kid_factor.reshape(-1)[:, None, None]
    j = j.expand(B, c.J, 3)
    pos = _positions(c, glob, j)
    tr = torch.zeros((1, 3), dtype=dt, device=dev) if trans is None else trans
    out = [pos + tr[:, None], glob]
    if not return_vertices:
        return out + [None]
    v = c.v_template[None] + torch.einsum('vcp,bp->bvc', c.posedirs, rel1.reshape(B, (c.J - 1) * 9))
    if nb:
        v = v + torch.einsum('vcp,bp->bvc', c.shapedirs[:, :, :nb], shape_betas[:, :nb])
    if kid_factor is not None:
        v = v + c.kid_shapedir[None] * kid_factor.reshape(-1)[:, None, None]
    transl = pos - _mv3(glob, j)
    blend = torch.einsum('vj,bjk->bvk', c.weights, torch.cat([glob.reshape(B, c.J, 9), transl], 2))
    verts = _mv3(blend[..., :9].reshape(B, c.V, 3, 3), v) + blend[..., 9:]  # (not einsum: B V batched 3x3 GEMVs)
    return out + [verts + tr[:, None]]


# ----------------------------------------------------------------------------------------------------------------------
# the fit (pt/bodyfitter.py:283-549 and the stages it calls)
# ----------------------------------------------------------------------------------------------------------------------
def _part_sums(c: Constants, t, a, vw):
    """Per-part sums over the used vertices (pt/bodyfitter.py:235-280): raw = sum t a_w^T, s_t, s_a, s_w."""
    B = t.shape[0]
    tu, au = t[:, c.used], a[:, c.used].expand(B, -1, -1)
    if vw is not None:
        wu = vw[:, c.used, None]
        au, ts = au * wu, tu * wu
    else:
        ts = tu
    z = lambda *s: torch.zeros(s, dtype=t.dtype, device=t.device)  # noqa: E731
    raw = z(B, c.J, 9).index_add(1, c.part_of_used, _outer(tu, au).reshape(B, -1, 9)).reshape(B, c.J, 3, 3)
    s_t = z(B, c.J, 3).index_add(1, c.part_of_used, ts)
    s_a = z(B, c.J, 3).index_add(1, c.part_of_used, au)
    if vw is not None:
        s_w = z(B, c.J).index_add(1, c.part_of_used, vw[:, c.used])
    else:
        s_w = c.part_counts[None].expand(B, c.J)
    return raw, s_t, s_a, s_w


def _fit_global_rotations(c: Constants, t, tj, a, aj, vw, jw):
    """Independent per-part rotations (pt/bodyfitter.py:1321-1416): Kabsch for leaf and multi-joint parts, swing from
    the bone direction + twist from the vertex covariance for two-joint parts."""
    if tj is None or aj is None:
        tj = torch.einsum('jv,bvc->bjc', c.J_regressor, t)
        aj = torch.einsum('jv,bvc->bjc', c.J_regressor, a)
    B = t.shape[0]
    aj = aj.expand(B, -1, -1)
    raw, s_t, s_a, s_w = _part_sums(c, t, a, vw)
    mt = torch.einsum('ij,bjc->bic', c.center, tj)
    ma = torch.einsum('ij,bjc->bic', c.center, aj)
    A = raw - _outer(s_t, ma) - _outer(mt, s_a) + s_w[..., None, None] * _outer(mt, ma)
    # multi-joint parts: covariance of the part's joints (self + children)
    rj, tsj = (aj, tj) if jw is None else (aj * jw[..., None], tj * jw[..., None])
    rawj = torch.einsum('mk,bkx,bky->bmxy', c.mjp, tj, rj)
    swj = c.mjp.sum(1)[None].expand(B, -1) if jw is None else torch.einsum('mk,bk->bm', c.mjp, jw)
    mtm, mam = mt[:, c.multi], ma[:, c.multi]
    Am = (rawj - _outer(torch.einsum('mk,bkx->bmx', c.mjp, tsj), mam) - _outer(mtm, torch.einsum('mk,bkx->bmx', c.mjp, rj))
          + swj[..., None, None] * _outer(mtm, mam))
    Rm = proj_so3(Am)
    Rl = proj_so3(A[:, c.leaf])
    # bone parts
    Ab = A[:, c.bone]
    j0, j1 = c.bone_pairs[:, 0], c.bone_pairs[:, 1]
    b_ref = _unit(aj[:, j1] - aj[:, j0])
    b_tgt = _unit(tj[:, j1] - tj[:, j0])
    Rs = _align_unit_vectors(b_ref, b_tgt)
    H = _mm3(Rs, Ab.transpose(-1, -2))
    trH = H[..., 0, 0] + H[..., 1, 1] + H[..., 2, 2]
    bHb = (b_tgt * _mv3(H, b_tgt)).sum(-1)
    vee = torch.stack([H[..., 1, 2] - H[..., 2, 1], H[..., 2, 0] - H[..., 0, 2], H[..., 0, 1] - H[..., 1, 0]], -1)
    ang = torch.atan2((b_tgt * vee).sum(-1), trH - bHb)
    Rb = _mm3(rotvec2mat(b_tgt * ang[..., None]), Rs)
    return torch.cat([Rm, Rl, Rb], 1)[:, c.assemble]


def _fit_shape(c: Constants, S: int, glob, t, tj, vw, jw, reg: float, reg2: float, beta_ref, enable_kid: bool = False,
               kid_reg: Optional[float] = None, kid_ref=None, scale_mode: int = 0, scale_reg: float = 0.0,
               share_beta: bool = False):
    """Shape (+ kid factor, + scale correction) and translation for given orientations: weighted, centred normal
    equations (pt/bodyfitter.py:863-1319; ``share_beta``: pt/lstsq.py:32-90).  ``scale_mode``: 0 none, 1 target,
    2 fit.  Returns a dict with shape_betas, trans, relative_orientations, joints, vertices (+ kid_factor, scale_corr)."""
    B, J = t.shape[0], c.J
    rel = _relative(c, glob)
    sd, Jt = c.shapedirs[:, :, :S], torch.cat([c.J_template[:, :, None], c.J_shapedirs[:, :, :S]], 2)
    if enable_kid:
        sd = torch.cat([sd, c.kid_shapedir[:, :, None]], 2)
        Jt = torch.cat([Jt, c.kid_J_shapedir[:, :, None]], 2)
    n_sh = sd.shape[2]  # shared-able unknowns: betas (+ kid)
    # positions of the joints as affine functions of the unknowns: root + path sums of the parent-rotated bone columns
    P = Jt[0][None, None] + torch.einsum('ik,bkCs->biCs', c.anc1, _mm3(glob[:, c.par1], (Jt[1:] - Jt[c.par1])[None]))
    T = P - _mm3(glob, Jt[None])
    v_posed = c.v_template[None] + torch.einsum('vcp,bp->bvc', c.posedirs, rel[:, 1:].reshape(B, (J - 1) * 9))
    blend = torch.einsum('vj,bjk->bvk', c.weights, glob.reshape(B, J, 9)).reshape(B, c.V, 3, 3)
    ext = torch.cat([_mv3(blend, v_posed)[..., None],
                     torch.einsum('bvCc,vcs->bvCs', blend, sd)], 3)
    ext = ext + torch.einsum('vj,bjCs->bvCs', c.weights, T)
    if tj is None:
        tgt, full = t, ext
    else:
        tgt, full = torch.cat([t, tj], 1), torch.cat([ext, P], 1)
    pos, jac = full[..., 0], full[..., 1:]
    if scale_mode == 1:
        jac = torch.cat([jac, -tgt[..., None]], 3)
    elif scale_mode == 2:
        jac = torch.cat([jac, pos[..., None]], 3)
    b = tgt - pos
    # weights enter the shape stage only as a complete set (pt/bodyfitter.py:1018-1028)
    if tj is not None and vw is not None and jw is not None:
        w = torch.cat([vw, jw], 1)
    elif tj is None and vw is not None:
        w = vw
    else:
        w = torch.ones(tgt.shape[:2], dtype=t.dtype, device=t.device)
    wsum = w.sum(1)
    ws = torch.where(wsum == 0, torch.ones_like(wsum), wsum)
    mean_A = torch.einsum('bn,bncs->bcs', w, jac) / ws[:, None, None]
    mean_b = torch.einsum('bn,bnc->bc', w, b) / ws[:, None]
    Ac, bc = jac - mean_A[:, None], b - mean_b[:, None]
    WA = Ac * w[:, :, None, None]
    d = torch.float64
    lam = [reg2] * min(2, S) + [reg] * max(S - 2, 0)
    ref = torch.zeros((1, S), dtype=d, device=t.device)
    if beta_ref is not None:
        n = min(beta_ref.shape[1], S)
        ref = torch.cat([beta_ref[:, :n].to(d), torch.zeros((beta_ref.shape[0], S - n), dtype=d, device=t.device)], 1)
    if enable_kid:
        lam.append(reg if kid_reg is None else kid_reg)
        kr = torch.zeros((1, 1), dtype=d, device=t.device) if kid_ref is None else kid_ref.reshape(-1, 1).to(d)
        ref = torch.cat([ref.expand(max(ref.shape[0], kr.shape[0]), -1), kr.expand(max(ref.shape[0], kr.shape[0]), -1)], 1)
    if scale_mode:
        lam.append(scale_reg)
        ref = torch.cat([ref, torch.zeros((ref.shape[0], 1), dtype=d, device=t.device)], 1)
    lam = torch.tensor(lam, dtype=d, device=t.device)
    G = torch.einsum('bncs,bnct->bst', WA, Ac).to(d) + torch.diag(lam)[None]
    r = torch.einsum('bncs,bnc->bs', WA, bc).to(d)
    npar = G.shape[-1]
    if share_beta and npar > n_sh:
        # betas (+ kid) shared over the batch, the scale column per instance: the per-instance unknown is eliminated
        # (Schur complement), the shared system is summed over the batch.  The regulariser enters lstsq_partial_share
        # as extra rows (weight lam_i, right-hand side lam_i ref_i), hence lam^2 ref here.
        rp = r + lam * lam * ref
        Gss, Gsz, Gzz = G[:, :n_sh, :n_sh], G[:, :n_sh, n_sh:], G[:, n_sh:, n_sh:]
        c_s = torch.linalg.solve(Gzz, Gsz.transpose(1, 2))
        c_r = torch.linalg.solve(Gzz, rp[:, n_sh:, None])
        x_s = torch.linalg.solve((Gss - Gsz @ c_s).sum(0), (rp[:, :n_sh, None] - Gsz @ c_r).sum(0))
        x_z = c_r - c_s @ x_s[None]
        x = torch.cat([x_s[None, :, 0].expand(B, -1), x_z[:, :, 0]], 1)
    elif share_beta:
        # lstsq(..., shared=True): the regulariser sits inside every instance's Gramian, no reference term
        x = torch.linalg.solve(G.sum(0), r.sum(0)[:, None])[None, :, 0].expand(B, -1)
    else:
        x = torch.linalg.solve(G, (r + lam * ref)[..., None])[..., 0]
    x = x.to(t.dtype)
    trans = mean_b - torch.einsum('bcs,bs->bc', mean_A, x)
    out = {'relative_orientations': rel}
    beta = x[:, :S]
    kid = x[:, S] if enable_kid else None
    out['shape_betas'], out['trans'] = beta, trans
    if kid is not None:
        out['kid_factor'] = kid
    if scale_mode:
        sc = x[:, -1] + 1
        out['scale_corr'] = sc
        if scale_mode == 2:  # the returned betas / kid factor stay unscaled; the mesh of the next stage uses beta / s
            beta = beta / sc[:, None]
            kid = kid / sc if kid is not None else None
    coef = beta if kid is None else torch.cat([beta, kid[:, None]], 1)
    out['joints'] = P[..., 0] + torch.einsum('bjcs,bs->bjc', P[..., 1:], coef) + trans[:, None]
    out['vertices'] = ext[..., 0] + torch.einsum('bvcs,bs->bvc', ext[..., 1:], coef) + trans[:, None]
    return out


def _fit_global_rotations_dependent(c: Constants, S: int, t, tj, a, aj, vw, jw, R_prev, betas, trans, kid=None,
                                    scale_corr=None):
    """Final adjustment along the kinematic chain (pt/bodyfitter.py:1418-1469, :1546-1595): each adjustable part is
    re-fitted about the position its joint gets from the already adjusted parents."""
    true_aj = aj
    if tj is None or aj is None:
        tj = torch.einsum('jv,bvc->bjc', c.J_regressor, t)
        aj = torch.einsum('jv,bvc->bjc', c.J_regressor, a)
    if true_aj is None:
        true_aj = aj
    p = c.plan
    j = c.J_template[None] + torch.einsum('jcs,bs->bjc', c.J_shapedirs[:, :, :S], betas[:, :S])
    if kid is not None:
        j = j + c.kid_J_shapedir[None] * kid[:, None, None]
    if scale_corr is not None:
        j = j * scale_corr[:, None, None]
    raw, s_t, s_a, s_w = _part_sums(c, t, a, vw)
    R: List[torch.Tensor] = [R_prev[:, i] for i in range(c.J)]
    pos: List[Optional[torch.Tensor]] = [None] * c.J
    adjustable = set(p.adjustable_parts)
    for i in range(c.J):
        if i == 0:
            pos[0] = j[:, 0] + trans
        else:
            q = c.parents[i]
            pos[i] = pos[q] + _mv3(R[q], j[:, i] - j[:, q])
        if p.is_smpl_family and i in (10, 11):
            R[i] = R[7 if i == 10 else 8]
            continue
        if i not in adjustable:
            continue
        c_t, c_a = pos[i], true_aj[:, i]
        A = raw[:, i] - _outer(s_t[:, i], c_a) - _outer(c_t, s_a[:, i]) + s_w[:, i, None, None] * _outer(c_t, c_a)
        cas = p.children_and_self[i]
        ej = tj[:, cas] - c_t[:, None]
        dj = aj[:, cas] - c_a[:, None]
        if jw is not None:
            dj = dj * jw[:, cas, None]
        A = A + (ej.unsqueeze(-1) * dj.unsqueeze(-2)).sum(1)
        R[i] = _mm3(proj_so3(A), R_prev[:, i])
    return torch.stack(R, 1)


def fit(bm, n_betas: int, target_vertices, target_joints=None, vertex_weights=None, joint_weights=None,
        num_iter: int = 1, beta_regularizer: float = 1.0, beta_regularizer2: float = 0.0,
        final_adjust_rots: bool = True, initial_pose_rotvecs=None, initial_shape_betas=None,
        want_pose_rotvecs: bool = True, want_rel_orient: bool = False, enable_kid: bool = False,
        scale_regularizer: float = 0.0, kid_regularizer: Optional[float] = None, share_beta: bool = False,
        scale_target: bool = False, scale_fit: bool = False, initial_kid_factor=None):
    """Differentiable evaluation of the closed-form fit (driver: pt/bodyfitter.py:283-549).  Returns (shape_betas,
    trans, orientations, relative_orientations, pose_rotvecs | None, kid_factor | None, scale_corr | None) -- the
    tensors of the ``smplfit_b200::fit`` op in its order."""
    t = target_vertices
    c = constants(bm, t.dtype, t.device)
    S = n_betas
    tj, vw, jw = target_joints, vertex_weights, joint_weights
    scale_mode = 1 if scale_target else (2 if scale_fit else 0)
    if tj is None:
        mean = t.mean(1)
        t = t - mean[:, None]
    else:
        mean = torch.cat([t, tj], 1).mean(1)
        t, tj = t - mean[:, None], tj - mean[:, None]
    if initial_pose_rotvecs is not None or initial_shape_betas is not None:
        ij, io, iv = lbs(c, pose_rotvecs=initial_pose_rotvecs, shape_betas=initial_shape_betas, kid_factor=initial_kid_factor)
        glob = _mm3(_fit_global_rotations(c, t, tj, iv, ij, vw, jw), io)
    else:
        if c.template_mesh is None:
            with torch.no_grad():
                c.template_mesh = lbs(c)[2]
        glob = _fit_global_rotations(c, t, tj, c.template_mesh, c.J_template[None], vw, jw)
    kid_ref = initial_kid_factor if enable_kid else None
    shape = lambda g, sm, sr: _fit_shape(c, S, g, t, tj, vw, jw, beta_regularizer, beta_regularizer2,  # noqa: E731
                                         initial_shape_betas, enable_kid, kid_regularizer, kid_ref, sm, sr, share_beta)
    for _ in range(num_iter - 1):
        res = shape(glob, 0, 0.0)
        glob = _mm3(_fit_global_rotations(c, t, tj, res['vertices'], res['joints'] if tj is not None else None, vw, jw), glob)
    res = shape(glob, scale_mode, scale_regularizer)
    betas, trans, rel = res['shape_betas'], res['trans'], res['relative_orientations']
    kid, sc = res.get('kid_factor'), res.get('scale_corr')
    if final_adjust_rots:
        rv, rj = res['vertices'], res['joints']
        if scale_mode == 1:
            s3 = sc[:, None, None]
            glob = _fit_global_rotations_dependent(c, S, t * s3, tj * s3 if tj is not None else None, rv, rj, vw, jw, glob,
                                                   betas, trans, kid)
        elif scale_mode == 2:
            s3, tr = sc[:, None, None], trans[:, None]
            glob = _fit_global_rotations_dependent(c, S, t, tj, s3 * rv + (1 - s3) * tr, s3 * rj + (1 - s3) * tr, vw, jw,
                                                   glob, betas, trans, kid, sc)
        else:
            glob = _fit_global_rotations_dependent(c, S, t, tj, rv, rj, vw, jw, glob, betas, trans, kid)
    if scale_mode == 1:
        trans = trans + mean * sc[:, None]
    elif scale_mode == 2:
        trans = trans + mean / sc[:, None]
    else:
        trans = trans + mean
    if want_pose_rotvecs or want_rel_orient:
        rel = _relative(c, glob)
    rotvecs = mat2rotvec(rel).reshape(glob.shape[0], -1) if want_pose_rotvecs else None
    return betas, trans, glob, rel, rotvecs, kid, sc


def _centre(t, tj):
    if tj is None:
        mean = t.mean(1)
        return t - mean[:, None], None, mean
    mean = torch.cat([t, tj], 1).mean(1)
    return t - mean[:, None], tj - mean[:, None], mean


def fit_with_known_pose(bm, n_betas: int, enable_kid: bool, pose_rotvecs, target_vertices, target_joints=None,
                        vertex_weights=None, joint_weights=None, beta_regularizer_reference=None,
                        kid_regularizer_reference=None, beta_regularizer: float = 1.0, beta_regularizer2: float = 0.0,
                        scale_regularizer: float = 0.0, kid_regularizer: Optional[float] = None, share_beta: bool = False,
                        scale_target: bool = False, scale_fit: bool = False):
    """Differentiable evaluation of ``BodyFitter.fit_with_known_pose`` (pt/bodyfitter.py:552-653): one shape stage with
    the orientations of the given pose.  Returns the result dict of the method."""
    c = constants(bm, target_vertices.dtype, target_vertices.device)
    t, tj, mean = _centre(target_vertices, target_joints)
    B = t.shape[0]
    glob = lbs(c, pose_rotvecs=pose_rotvecs, return_vertices=False)[1].expand(B, -1, -1, -1)
    res = _fit_shape(c, n_betas, glob, t, tj, vertex_weights, joint_weights, beta_regularizer, beta_regularizer2,
                     beta_regularizer_reference, enable_kid, kid_regularizer,
                     kid_regularizer_reference if enable_kid else None, 1 if scale_target else (2 if scale_fit else 0),
                     scale_regularizer, share_beta)
    out = {'shape_betas': res['shape_betas'], 'trans': res['trans'] + mean, 'relative_orientations': res['relative_orientations']}
    for k in ('kid_factor', 'scale_corr'):
        if k in res:
            out[k] = res[k]
    return out


def _fit_scale_and_translation(t, a, tj, aj, vw, jw, scale: bool):
    """pt/bodyfitter.py:1628-1681: weighted centroids (and the ratio of the weighted spreads)."""
    if tj is None or aj is None:
        tb, ab = t, a
        w = vw if vw is not None else torch.ones(t.shape[:2], dtype=t.dtype, device=t.device)
    else:
        tb, ab = torch.cat([t, tj], 1), torch.cat([a, aj], 1)
        w = torch.cat([vw, jw], 1) if (vw is not None and jw is not None) else torch.ones(tb.shape[:2], dtype=t.dtype,
                                                                                         device=t.device)
    w = (w / w.sum(1, keepdim=True))[..., None]
    mt, ma = (tb * w).sum(1), (ab * w).sum(1)
    if not scale:
        return None, mt - ma
    sc = ((((tb - mt[:, None]) ** 2) * w).sum((1, 2)) / (((ab - ma[:, None]) ** 2) * w).sum((1, 2))).sqrt()
    return sc, mt - sc[:, None] * ma


def fit_with_known_shape(bm, n_betas: int, shape_betas, target_vertices, target_joints=None, vertex_weights=None,
                         joint_weights=None, kid_factor=None, initial_pose_rotvecs=None, num_iter: int = 1,
                         final_adjust_rots: bool = True, scale_fit: bool = False, want_pose_rotvecs: bool = True,
                         want_rel_orient: bool = False):
    """Differentiable evaluation of ``BodyFitter.fit_with_known_shape`` (pt/bodyfitter.py:656-838): rotation fits
    against the forward pass of the known betas, then scale / translation, then the final adjustment."""
    c = constants(bm, target_vertices.dtype, target_vertices.device)
    t, tj, mean = _centre(target_vertices, target_joints)
    vw, jw = vertex_weights, joint_weights
    B = t.shape[0]
    betas = shape_betas.expand(B, -1)
    kid = None if kid_factor is None else kid_factor.reshape(-1).expand(B)
    ij, io, iv = lbs(c, pose_rotvecs=initial_pose_rotvecs, shape_betas=betas, kid_factor=kid)
    glob = _mm3(_fit_global_rotations(c, t, tj, iv, ij, vw, jw), io)
    for _ in range(num_iter - 1):
        rj, _, rv = lbs(c, glob_rotmats=glob, shape_betas=betas, kid_factor=kid)
        glob = _mm3(_fit_global_rotations(c, t, tj, rv, rj if tj is not None else None, vw, jw), glob)
    rj, _, rv = lbs(c, glob_rotmats=glob, shape_betas=betas, kid_factor=kid)
    sc, tr = _fit_scale_and_translation(t, rv, tj, rj, vw, jw, scale_fit)
    if final_adjust_rots:
        nb = min(betas.shape[1], c.S)
        if scale_fit:
            s3 = sc[:, None, None]
            glob = _fit_global_rotations_dependent(c, nb, t, tj, s3 * rv + tr[:, None], s3 * rj + tr[:, None], vw, jw, glob,
                                                   betas, tr, kid, sc)
        else:
            glob = _fit_global_rotations_dependent(c, nb, t, tj, rv + tr[:, None], rj + tr[:, None], vw, jw, glob, betas,
                                                   tr, kid)
    out = {'trans': tr + mean, 'orientations': glob}
    if scale_fit:
        out['scale_corr'] = sc
    if want_pose_rotvecs or want_rel_orient:
        out['relative_orientations'] = _relative(c, glob)
        if want_pose_rotvecs:
            out['pose_rotvecs'] = mat2rotvec(out['relative_orientations']).reshape(B, -1)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# backward of the custom ops: slice the batch, re-evaluate with grad, pull the cotangents back
# ----------------------------------------------------------------------------------------------------------------------
def _slices(B: int, per_instance_bytes: float, budget: Optional[float] = None, device=None, cap: int = 512):
    """Cut the batch so that one slice's saved intermediates stay within ``budget`` bytes (default: 30 % of the free
    device memory, at least 1.5 GB) and within ``cap`` instances (beyond a few hundred instances the evaluation is
    no longer launch-bound, so larger slices only cost memory)."""
    if budget is None:
        budget = 1.5e9
        if device is not None and torch.device(device).type == 'cuda':
            budget = max(budget, 0.3 * torch.cuda.mem_get_info(device)[0])
    n = int(max(1, min(B, cap, budget // max(per_instance_bytes, 1.0))))
    return [(a, min(a + n, B)) for a in range(0, B, n)]


def _pullback(B: int, inputs, needs, run, grads_out, slices, device):
    """``inputs``: tensors or None; ``needs[i]``: gradient wanted; ``run(*sliced_inputs)`` -> outputs (None allowed);
    ``grads_out``: cotangents aligned with the outputs (None / empty allowed).  Inputs with batch 1 (broadcast) get
    the sum over the slices.  The evaluation runs in float32 on ``device`` (the model's); the gradients are returned in
    the dtype and on the device of their inputs."""
    grads = [torch.zeros(x.shape, dtype=torch.float32, device=device) if (x is not None and nd) else None
             for x, nd in zip(inputs, needs)]
    for a, b in slices:
        xs = []
        for x, nd in zip(inputs, needs):
            if x is None:
                xs.append(None)
                continue
            xi = x[a:b] if x.shape[0] == B else x
            xi = xi.detach().to(device=device, dtype=torch.float32)
            xs.append(xi.requires_grad_(True) if nd else xi)
        with torch.enable_grad():
            outs = run(*xs)
            pairs = [(o, g[a:b].to(device)) for o, g in zip(outs, grads_out)
                     if o is not None and g is not None and g.numel() > 0 and o.requires_grad]
            wanted = [x for x, nd in zip(xs, needs) if nd]
            if not pairs or not wanted:
                continue
            got = torch.autograd.grad([o for o, _ in pairs], wanted, [g.to(o.dtype) for o, g in pairs],
                                      allow_unused=True)
        k = 0
        for i, nd in enumerate(needs):
            if not nd:
                continue
            g = got[k]
            k += 1
            if g is None:
                continue
            if inputs[i].shape[0] == B:
                grads[i][a:b] += g
            else:
                grads[i] += g
    return [g if g is None else g.to(device=x.device, dtype=x.dtype) for g, x in zip(grads, inputs)]


def forward_backward(bm, tensors, needs, return_vertices: bool, grads_out):
    """Backward of ``smplfit_b200::forward``.  ``tensors`` = (pose_rotvecs, shape_betas, trans, kid_factor, rel_rotmats,
    glob_rotmats); ``grads_out`` = cotangents of (joints, orientations, vertices)."""
    B = 1
    for x in (tensors[0], tensors[1], tensors[2], tensors[4], tensors[5]):
        if x is not None:
            B = x.shape[0]
            break
    dev = bm.v_template.device
    c = constants(bm, torch.float32, dev)

    def run(pose, betas, trans, kid, rel, glob):
        return lbs(c, pose, betas, trans, kid, rel, glob, return_vertices)

    per = 4.0 * c.V * (3 * 12 + 9 + c.J)  # posed template, blended transforms, outputs
    return _pullback(B, tensors, needs, run, grads_out, _slices(B, per, device=dev, cap=4096), dev)


def fit_backward(fitter, tensors, needs, opts: dict, grads_out):
    """Backward of ``smplfit_b200::fit``.  ``tensors`` = (target_vertices, target_joints, vertex_weights, joint_weights,
    initial_pose_rotvecs, initial_shape_betas, initial_kid_factor); ``grads_out`` = cotangents of (shape_betas, trans,
    orientations, relative_orientations, pose_rotvecs, kid_factor, scale_corr); ``opts`` = the keyword options of
    ``fit`` above.  ``share_beta`` couples the instances: one slice."""
    bm = fitter.body_model
    B = tensors[0].shape[0]
    S = fitter.n_betas

    def run(tv, tj, vw, jw, ip, ib, ik):
        return fit(bm, S, tv, tj, vw, jw, initial_pose_rotvecs=ip, initial_shape_betas=ib, initial_kid_factor=ik,
                   enable_kid=fitter.enable_kid, **opts)

    # saved per shape stage: the (V, 3, 1 + S) skinned basis, its centred / weighted copies, the blend matrices
    # (measured on the B200: 12 MB per SMPL instance at num_iter = 3, this estimate gives 15 MB)
    per = 4.0 * bm.num_vertices * 3 * (S + 14) * 2.5 * max(1, opts['num_iter'])
    dev = bm.v_template.device
    slices = [(0, B)] if opts.get('share_beta') else _slices(B, per, device=dev)
    return _pullback(B, tensors, needs, run, grads_out, slices, dev)


class _Recompute(torch.autograd.Function):
    """Values from ``run_cuda`` (the CUDA entry point, no grad), gradient by re-evaluating ``run_torch`` on slices of
    the batch: the wrapper of the methods that do not go through a custom op (``fit_with_known_pose`` /
    ``fit_with_known_shape``).  Both callables take the tensor arguments positionally and return the result dict."""

    @staticmethod
    def forward(ctx, run_cuda, run_torch, keys, B, device, per, coupled, *tensors):
        with torch.no_grad():
            res = run_cuda(*[None if x is None else x.detach() for x in tensors])
        ctx.run_torch, ctx.keys, ctx.B, ctx.device, ctx.per, ctx.coupled = run_torch, keys, B, device, per, coupled
        ctx.needs = [x is not None and x.requires_grad for x in tensors]
        ctx.save_for_backward(*tensors)
        return tuple(res[k] for k in keys)

    @staticmethod
    def backward(ctx, *grads):
        tensors = list(ctx.saved_tensors)
        keys = ctx.keys

        def run(*xs):
            res = ctx.run_torch(*xs)
            return [res[k] for k in keys]

        slices = [(0, ctx.B)] if ctx.coupled else _slices(ctx.B, ctx.per, device=ctx.device)
        g = _pullback(ctx.B, tensors, ctx.needs, run, list(grads), slices, ctx.device)
        return (None,) * 7 + tuple(g)


def differentiable_call(run_cuda, run_torch, keys, B: int, device, per_instance_bytes: float, coupled: bool, tensors):
    """-> result dict whose tensors carry the gradient of ``run_torch`` and the values of ``run_cuda``."""
    outs = _Recompute.apply(run_cuda, run_torch, tuple(keys), B, device, per_instance_bytes, coupled, *tensors)
    return dict(zip(keys, outs))
