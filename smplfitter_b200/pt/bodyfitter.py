"""``BodyFitter`` -- drop-in for ``smplfitter.pt.BodyFitter``
(/root/reference/src/smplfitter/pt/bodyfitter.py:15-1681): closed-form inverse of the body
model.  ``fit`` / ``fit_with_known_pose`` run entirely in the sm_100a CUDA library
(``smplfit_fit`` / ``smplfit_fit_known_pose``); this module only validates arguments, owns the
static index tables (bit-exact with the reference's ``__init__``, see masks.py) and marshals
pointers.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _native
from . import _ops


class BodyFitter(_ops.RegisteredModule, nn.Module):
    """Fits body-model parameters to target vertices (and optionally joints).

    Parameters:
        body_model: the ``smplfitter_b200.pt.BodyModel`` to fit.
        enable_kid: adds the kid blend-shape factor as an unknown (pt/bodyfitter.py:25-33).
    """

    def __init__(self, body_model, enable_kid: bool = False):
        super().__init__()
        self.body_model = body_model
        self.n_betas = body_model.shapedirs.shape[2]
        self.enable_kid = enable_kid
        plan = body_model._plan
        self.is_smpl_family = plan.is_smpl_family
        # reference-named static tables (pt/bodyfitter.py:36-233)
        i64 = lambda x: torch.from_numpy(np.array(x, dtype=np.int64, order='C'))  # noqa: E731
        self.part_assignment = nn.Buffer(i64(plan.part_assignment))
        self.part_vertex_selectors = [i64(s) for s in plan.part_vertex_selectors]
        self.children_and_self = plan.children_and_self
        self.descendants_and_self = plan.descendants_and_self
        self.multi_joint_parts = plan.multi_joint_parts
        self.bone_parts = plan.bone_parts
        self.leaf_parts = plan.leaf_parts
        self.adjustable_parts = plan.adjustable_parts
        self.used_vertex_indices = nn.Buffer(i64(plan.used_vertex_indices))
        self.assemble_indices = nn.Buffer(i64(plan.assemble_indices))
        self.bone_pairs = nn.Buffer(i64(plan.bone_pairs))
        self.fk_js = nn.Buffer(i64(plan.fk_js))
        self.fk_ps = nn.Buffer(i64(plan.fk_ps))
        self.fk_level_sizes = plan.fk_level_sizes
        self.num_fk_levels = len(plan.fk_level_sizes)
        self.adj_parts = nn.Buffer(i64(plan.adj_parts))
        self.adj_level_sizes = plan.adj_level_sizes
        self.adj_last_level = plan.adj_last_level
        self.adj_part_joints = nn.Buffer(i64(plan.adj_part_joints))
        self.leveladj_supported = plan.leveladj_supported
        self.cas_flat = nn.Buffer(i64(plan.cas_flat))
        self.cas_starts = plan.cas_starts
        self.default_mesh_tf = body_model._t_template_mesh
        J, V, S = body_model.num_joints, body_model.num_vertices, self.n_betas
        self.gram_supported = (not enable_kid) and J * 3 * V * S <= 2 ** 26
        # unknown-wise extended tables: [betas | kid]
        sd = body_model.shapedirs
        jt = [body_model.J_template.reshape(-1, 3, 1), body_model.J_shapedirs]
        if enable_kid:
            sd = torch.cat([sd, body_model.kid_shapedir[:, :, None]], dim=2)
            jt.append(body_model.kid_J_shapedir.reshape(-1, 3, 1))
        self.register_buffer('_t_fit_shapedirs', sd.contiguous().clone(), persistent=False)
        self.J_template_ext = nn.Buffer(torch.cat(jt, dim=2).contiguous())
        self._ns = S + (1 if enable_kid else 0)
        if self._ns > 17:
            raise NotImplementedError(
                f'smplfitter_b200.BodyFitter solves for at most 17 shape unknowns (betas + kid); this body model has '
                f'{S} betas{" + kid" if enable_kid else ""}. Construct the BodyModel with num_betas<=16 '
                f"(e.g. BodyModel('{body_model.model_name}', num_betas=16)); the reference solves all {S}, so a "
                'truncated model is a different (smaller) least-squares problem, not an approximation made silently.')
        # packed per-vertex records (internal order) and the closed-form SA constants
        ns = self._ns
        order = body_model._t_order.cpu().numpy()
        sd_np = sd.cpu().numpy().astype(np.float32)
        w_np = body_model.weights.cpu().numpy()
        wS = np.einsum('vk,vcs->kcs', w_np.astype(np.float64), sd_np.astype(np.float64))
        self.register_buffer('_t_fit_wS', torch.from_numpy(np.ascontiguousarray(wS)), persistent=False)
        self.register_buffer('_t_fit_wsum', torch.tensor(w_np.astype(np.float64).sum(axis=0)), persistent=False)
        K = body_model._dims['skin_k']
        nsp = (ns + 1) // 2 * 2
        self._rec_len = (8 + 3 * nsp + 3) // 4 * 4
        if K <= 4:
            idx = body_model._t_skin_idx.cpu().numpy()
            ww = body_model._t_skin_w.cpu().numpy()
            idx4 = np.zeros((V, 4), np.int32)
            w4 = np.zeros((V, 4), np.float32)
            idx4[:, :K], w4[:, :K] = idx, ww
            idx4[:, K:] = idx[:, :1]
            srt = np.argsort(-w4, axis=1, kind='stable')
            idx4 = np.take_along_axis(idx4, srt, axis=1)
            w4 = np.take_along_axis(w4, srt, axis=1)
            rec = np.zeros((V, self._rec_len), np.float32)
            rec[:, 0:4] = w4
            rec[:, 4:8] = idx4.view(np.float32)
            for x in range(3):
                rec[:, 8 + x * nsp:8 + x * nsp + ns] = sd_np[:, x, :]
            self.register_buffer('_t_fit_rec', torch.from_numpy(np.ascontiguousarray(rec[order])), persistent=False)
            # per vertex (internal order): bit k set when skinning slot k must (re)load its joint rows, i.e. its weight
            # is non-zero and its joint differs from the last one loaded for that slot in the same segment -- the
            # kernels' register cache of joint rows (lite_kernels.cuh JointCache) replayed on the host
            seg_start = body_model._t_seg_start.cpu().numpy()
            idx_o, w_o = idx4[order], w4[order]
            mask = np.zeros(V, np.uint8)
            for a, b in zip(seg_start[:-1], seg_start[1:]):
                cached = [-1, -1, -1, -1]
                for i in range(int(a), int(b)):
                    m = 0
                    for k in range(4):
                        if w_o[i, k] != 0 and idx_o[i, k] != cached[k]:
                            cached[k] = int(idx_o[i, k])
                            m |= 1 << k
                    mask[i] = m
            self.register_buffer('_t_fit_slot_mask', torch.tensor(mask), persistent=False)
            self._build_pair_constants(idx4, w4, sd_np, J, ns)
            self._build_fused_tables(idx_o, w_o, sd_np[order], order, seg_start, ns)
        else:
            self._t_fit_rec = None
            self._gcf_npairs = 0
            self._fq = None
        self._handle = _ops.register(self)
        self.to(body_model.v_template.device)

    @torch.jit.unused
    def _build_pair_constants(self, idx4, w4, sd, J, ns):
        """Model constants of the closed-form Gramian (include/smplfit_b200.h, ``gcf_*``), in float64:
        over the joint pairs (k, l) sharing a vertex, A_kl[a,b,s,t] = sum_v w_vk w_vl S_vs[a] S_vt[b],
        Bm_kl[a,s] = sum_v w_vk w_vl S_vs[a], W_kl = sum_v w_vk w_vl."""
        V = idx4.shape[0]
        iu, ju = np.triu_indices(ns)
        ng = len(iu)
        ngp = (ng + 3) // 4 * 4
        # per pair: vertex lists and weight products (a vertex lists each joint once; padded slots have w = 0)
        prods = {}
        for a in range(4):
            for b in range(4):
                ww = w4[:, a].astype(np.float64) * w4[:, b].astype(np.float64)
                nz = np.nonzero(ww)[0]
                ka, kb = idx4[nz, a], idx4[nz, b]
                for k, l, v, x in zip(ka.tolist(), kb.tolist(), nz.tolist(), ww[nz].tolist()):
                    if k <= l and not (k == l and a != b):
                        prods.setdefault((k, l), []).append((v, x))
        S = sd.astype(np.float64)  # (V,3,ns)
        off = sorted(p for p in prods if p[0] != p[1])
        A = np.zeros((max(len(off), 1), 9, ngp), np.float32)
        G0 = np.zeros(ng, np.float64)
        Bm, Wkl = {}, {}
        for p, lst in prods.items():
            v = np.array([q[0] for q in lst])
            x = np.array([q[1] for q in lst])
            Sv = S[v]
            Bm[p] = np.einsum('v,vas->as', x, Sv)
            Wkl[p] = x.sum()
            full = np.einsum('v,vae,vbe->abe', x, Sv[:, :, iu], Sv[:, :, ju])  # A[a,b,(s,t)]
            if p[0] == p[1]:
                G0 += np.einsum('aae->e', full)
            else:
                A[off.index(p), :, :ng] = (full + full.transpose(1, 0, 2)).reshape(9, ng)
        lstart, lk, bm_cells, wh_cells = [0], [], [], []
        for l in range(J):
            for k in range(J):
                p = (min(k, l), max(k, l))
                if p in prods:
                    lk.append(k)
                    cell = np.zeros((3, (ns + 3) // 4 * 4))
                    cell[:, :ns] = Bm[p]
                    bm_cells.append(cell)
                    wh_cells.append(0.5 * Wkl[p])
            lstart.append(len(lk))
        i32 = lambda x: torch.from_numpy(np.array(x, dtype=np.int32, order='C'))  # noqa: E731
        reg = lambda n, t: self.register_buffer(n, t, persistent=False)  # noqa: E731
        reg('_t_gcf_pairs', i32(np.array(off or [(0, 0)], np.int32).reshape(-1, 2)))
        reg('_t_gcf_A', torch.from_numpy(np.ascontiguousarray(A)))
        # the same constants as a GEMM operand: AT[e][p*9 + ab], rows padded to the 256-row tile, K to 32, hi/lo split
        kt = (9 * max(len(off), 1) + 31) // 32 * 32
        AT = np.zeros(((ng + 255) // 256 * 256, kt), np.float32)
        AT[:ng, :9 * A.shape[0]] = A[:, :, :ng].reshape(-1, ng).T
        at_hi = (AT.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        reg('_t_gcf_AT_hi', torch.from_numpy(np.ascontiguousarray(at_hi)))
        reg('_t_gcf_AT_lo', torch.from_numpy(np.ascontiguousarray(AT - at_hi)))
        reg('_t_gcf_G0', torch.tensor(G0))
        reg('_t_gcf_lstart', i32(np.array(lstart)))
        reg('_t_gcf_lk', i32(np.array(lk)))
        reg('_t_gcf_Bm', torch.from_numpy(np.ascontiguousarray(np.stack(bm_cells).astype(np.float32))))
        reg('_t_gcf_Wh', torch.tensor(np.array(wh_cells, np.float32)))
        self._gcf_npairs = len(off)

    @torch.jit.unused
    def _build_fused_tables(self, idx_o, w_o, sd_o, order, seg_start, ns):
        """Constants of the fused fit passes (include/smplfit_b200.h ``fq_*``; csrc/fit_fused.cu): segment q owns the
        padded vertex slots [32 q, 32 q + 32); the GEMM constants are [posedirs | shapedirs (+ kid)] rows in slot order,
        scaled into fp16's normal range and split hi / lo; the slot records carry the skinning weights, the joints of the
        four register-cached slots and the bits saying which slots reload -- replayed here along the chain of segments
        one epilogue warp walks (every second segment)."""
        bm = self.body_model
        J, V = bm.num_joints, bm.num_vertices
        if J > 64 or int(np.max(np.diff(seg_start))) > 32:
            self._fq = None
            return
        P = 9 * (J - 1)
        nseg = len(seg_start) - 1
        nseg_pad = (nseg + 1) // 2 * 2
        slot_of = np.full(nseg_pad * 32, -1, np.int64)  # slot -> internal vertex position
        for q in range(nseg):
            a, b = int(seg_start[q]), int(seg_start[q + 1])
            slot_of[32 * q:32 * q + (b - a)] = np.arange(a, b)
        live = slot_of >= 0
        eye_feat = np.tile(np.eye(3), [J - 1, 1]).reshape(-1)
        pd = bm.posedirs.cpu().numpy().astype(np.float64)[order]  # (V,3,P) internal order
        v_rest = bm.v_template.cpu().numpy().astype(np.float64)[order] + np.einsum('vcp,p->vc', pd, eye_feat)
        rec = np.zeros((nseg_pad * 32, 8), np.uint32)
        for h in range(2):
            cached = [-1, -1, -1, -1]
            for q in range(h, nseg, 2):
                for i in range(int(seg_start[q]), int(seg_start[q + 1])):
                    m = 0
                    for k in range(4):
                        if w_o[i, k] != 0 and idx_o[i, k] != cached[k]:
                            cached[k] = int(idx_o[i, k])
                            m |= 1 << k
                    pack = (m << 24) | (1 << 28)
                    for k in range(4):
                        pack |= (max(cached[k], 0) & 63) << (6 * k)
                    rec[32 * q + (i - int(seg_start[q])), 4] = pack
        rec[live, 0:4] = np.ascontiguousarray(w_o, dtype=np.float32).view(np.uint32)[slot_of[live]]
        rec[live, 5:8] = v_rest.astype(np.float32).view(np.uint32)[slot_of[live]]
        # bits 29 / 30 of the first record of every 8-slot block: some slot reloads within the first / second four slots
        # of the block (the kernel's branch-free fast path needs four vertices without reloads)
        any_rl = ((rec[:, 4] >> 24) & 0xF).reshape(-1, 2, 4).max(axis=2) != 0
        rec[::8, 4] |= (any_rl[:, 0].astype(np.uint32) << 29) | (any_rl[:, 1].astype(np.uint32) << 30)
        nsp4 = (ns + 3) // 4 * 4  # rows padded to 16-byte multiples
        sdl = 3 * nsp4
        sd = np.zeros((nseg_pad * 32, sdl), np.float32)
        for x in range(3):
            sd[live, x * nsp4:x * nsp4 + ns] = sd_o[slot_of[live], x, :]
        kf = (P + ns + 31) // 32 * 32
        pm = np.zeros((nseg_pad * 32, 3, kf), np.float64)
        pm[live, :, :P] = pd[slot_of[live]]
        pm[live, :, P:P + ns] = sd_o[slot_of[live]].astype(np.float64)
        pm = pm.reshape(nseg_pad * 96, kf)
        amax = float(np.abs(pm).max())
        scale_log2 = int(np.clip(np.floor(np.log2(16384.0 / amax)) if amax > 0 else 0, 0, 14))
        pm *= 2.0 ** scale_log2
        hi = pm.astype(np.float16)
        lo = (pm - hi.astype(np.float64)).astype(np.float16)
        reg = lambda n, t: self.register_buffer(n, t, persistent=False)  # noqa: E731
        reg('_t_fq_P_hi', torch.from_numpy(hi))
        reg('_t_fq_P_lo', torch.from_numpy(lo))
        reg('_t_fq_rec', torch.from_numpy(rec.view(np.int32)))
        reg('_t_fq_sd', torch.from_numpy(sd))
        self._fq = dict(fq_kf=kf, fq_scale_log2=scale_log2, fq_sdl=sdl, fq_nseg_pad=nseg_pad)

    @torch.jit.unused
    def _struct(self) -> _native.ModelStruct:
        gcf = {}
        if self._gcf_npairs > 0:
            gcf = dict(gcf_npairs=self._gcf_npairs,
                       **{n: getattr(self, '_t_' + n).data_ptr() for n in
                          ('gcf_pairs', 'gcf_A', 'gcf_G0', 'gcf_lstart', 'gcf_lk', 'gcf_Bm', 'gcf_Wh', 'gcf_AT_hi',
                           'gcf_AT_lo')})
        fq = {}
        if self._fq is not None:
            fq = dict(**self._fq, **{n: getattr(self, '_t_' + n).data_ptr() for n in ('fq_P_hi', 'fq_P_lo', 'fq_rec', 'fq_sd')})
        return self.body_model._struct(dict(
            **gcf, **fq,
            fit_ns=self._ns,
            fit_shapedirs=self._t_fit_shapedirs.data_ptr(),
            fit_Jt_ext=self.J_template_ext.data_ptr(),
            fit_rec=0 if self._t_fit_rec is None else self._t_fit_rec.data_ptr(),
            fit_slot_mask=0 if self._t_fit_rec is None else self._t_fit_slot_mask.data_ptr(),
            fit_rec_len=self._rec_len,
            fit_wS=self._t_fit_wS.data_ptr(),
            fit_wsum=self._t_fit_wsum.data_ptr(),
        ))

    # ------------------------------------------------------------------------------
    @torch.jit.unused
    def _prep(self, x, shape, name):
        if x is None:
            return None
        if isinstance(x, np.ndarray):
            raise TypeError(f"Expected torch.Tensor for '{name}', got numpy.ndarray.")
        dev = self.body_model.v_template.device
        x = x.to(device=dev, dtype=torch.float32)
        if tuple(x.shape) != tuple(shape):
            raise ValueError(f"'{name}' must have shape {tuple(shape)}, got {tuple(x.shape)}")
        return x.contiguous()

    @torch.jit.unused
    def _opts(self, num_iter, final_adjust_rots, requested_keys, shape_weights, beta_regularizer,
              beta_regularizer2, kid_regularizer, scale_mode, scale_regularizer) -> _native.FitOpts:
        o = _native.FitOpts()
        o.num_iter = int(num_iter)
        o.final_adjust_rots = int(bool(final_adjust_rots))
        o.enable_kid = int(self.enable_kid)
        o.want_pose_rotvecs = int('pose_rotvecs' in requested_keys)
        o.want_rel_orient = int('relative_orientations' in requested_keys)
        o.shape_weights = int(shape_weights)
        o.scale_mode = scale_mode
        o.beta_regularizer = float(beta_regularizer)
        o.beta_regularizer2 = float(beta_regularizer2)
        o.kid_regularizer = float(beta_regularizer if kid_regularizer is None else kid_regularizer)
        o.scale_regularizer = float(scale_regularizer)
        return o

    @staticmethod
    @torch.jit.unused
    def _shape_weights_rule(target_joints, vertex_weights, joint_weights) -> bool:
        """pt/bodyfitter.py:1018-1028: weights enter the shape solve only as a complete set."""
        if target_joints is not None:
            return vertex_weights is not None and joint_weights is not None
        return vertex_weights is not None

    @torch.jit.unused
    def _pad_ref(self, ref, B, name):
        if ref is None:
            return None
        dev = self.body_model.v_template.device
        ref = ref.to(device=dev, dtype=torch.float32)
        out = torch.zeros((B, self.n_betas), device=dev, dtype=torch.float32)
        n = min(ref.shape[1], self.n_betas)
        out[:, :n] = ref[:, :n]
        return out

    # ------------------------------------------------------------------------------
    @torch.jit.export
    def fit(
        self,
        target_vertices: torch.Tensor,
        target_joints: Optional[torch.Tensor] = None,
        vertex_weights: Optional[torch.Tensor] = None,
        joint_weights: Optional[torch.Tensor] = None,
        num_iter: int = 1,
        beta_regularizer: float = 1.0,
        beta_regularizer2: float = 0.0,
        scale_regularizer: float = 0.0,
        kid_regularizer: Optional[float] = None,
        share_beta: bool = False,
        final_adjust_rots: bool = True,
        scale_target: bool = False,
        scale_fit: bool = False,
        initial_pose_rotvecs: Optional[torch.Tensor] = None,
        initial_shape_betas: Optional[torch.Tensor] = None,
        initial_kid_factor: Optional[torch.Tensor] = None,
        requested_keys: Optional[List[str]] = None,
    ) -> Dict[str, torch.Tensor]:
        """Fit pose, shape and translation (pt/bodyfitter.py:283-549; same arguments and result keys).

        TorchScript-compatible: dispatches through the ``smplfit_b200::fit`` custom op."""
        want_rv = True
        want_rel = False
        if requested_keys is not None:
            want_rv = 'pose_rotvecs' in requested_keys
            want_rel = 'relative_orientations' in requested_keys
        if scale_target and scale_fit:
            raise ValueError('Only one of estim_scale_target and estim_scale_fit can be True')
        kid_reg = float('nan')
        if kid_regularizer is not None:
            kid_reg = float(kid_regularizer)
        outs = torch.ops.smplfit_b200.fit(
            self._handle, target_vertices, target_joints, vertex_weights, joint_weights, num_iter,
            float(beta_regularizer), float(beta_regularizer2), float(scale_regularizer), kid_reg, share_beta,
            final_adjust_rots, scale_target, scale_fit, initial_pose_rotvecs, initial_shape_betas,
            initial_kid_factor, want_rv, want_rel)
        result: Dict[str, torch.Tensor] = {
            'shape_betas': outs[0], 'trans': outs[1], 'orientations': outs[2], 'relative_orientations': outs[3]}
        if want_rv:
            result['pose_rotvecs'] = outs[4]
        if self.enable_kid:
            result['kid_factor'] = outs[5]
        if scale_target or scale_fit:
            result['scale_corr'] = outs[6]
        return result

    @torch.jit.unused
    def _fit_impl(
        self,
        target_vertices: torch.Tensor,
        target_joints: Optional[torch.Tensor] = None,
        vertex_weights: Optional[torch.Tensor] = None,
        joint_weights: Optional[torch.Tensor] = None,
        num_iter: int = 1,
        beta_regularizer: float = 1,
        beta_regularizer2: float = 0,
        scale_regularizer: float = 0,
        kid_regularizer: Optional[float] = None,
        share_beta: bool = False,
        final_adjust_rots: bool = True,
        scale_target: bool = False,
        scale_fit: bool = False,
        initial_pose_rotvecs: Optional[torch.Tensor] = None,
        initial_shape_betas: Optional[torch.Tensor] = None,
        initial_kid_factor: Optional[torch.Tensor] = None,
        requested_keys: Optional[list] = None,
    ) -> Dict[str, torch.Tensor]:
        if requested_keys is None:
            requested_keys = ['pose_rotvecs']
        if scale_target and scale_fit:
            raise ValueError('Only one of estim_scale_target and estim_scale_fit can be True')
        scale_mode = 1 if scale_target else (2 if scale_fit else 0)
        bm = self.body_model
        dev = bm.v_template.device
        _native.require_cuda(bm.v_template, 'the body model')
        if isinstance(target_vertices, np.ndarray):
            raise TypeError("Expected torch.Tensor for 'target_vertices', got numpy.ndarray.")
        if target_vertices.ndim != 3:
            raise ValueError(f'target_vertices must be (batch, {bm.num_vertices}, 3)')
        B, V, J, S = target_vertices.shape[0], bm.num_vertices, bm.num_joints, self.n_betas
        tv = self._prep(target_vertices, (B, V, 3), 'target_vertices')
        tj = self._prep(target_joints, (B, J, 3), 'target_joints')
        vw = self._prep(vertex_weights, (B, V), 'vertex_weights')
        jw = self._prep(joint_weights, (B, J), 'joint_weights')
        new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)  # noqa: E731
        out = dict(shape_betas=new(B, S), trans=new(B, 3), orientations=new(B, J, 3, 3),
                   relative_orientations=new(B, J, 3, 3))
        rotvecs = new(B, 3 * J) if 'pose_rotvecs' in requested_keys else None
        kid = new(B) if self.enable_kid else None
        scale_corr = new(B) if scale_mode else None
        if B == 0:
            if rotvecs is not None:
                out['pose_rotvecs'] = rotvecs
            if kid is not None:
                out['kid_factor'] = kid
            if scale_corr is not None:
                out['scale_corr'] = scale_corr
            return out
        init_v = init_j = init_o = None
        if initial_pose_rotvecs is not None or initial_shape_betas is not None:
            init = bm(shape_betas=initial_shape_betas, kid_factor=initial_kid_factor,
                      pose_rotvecs=initial_pose_rotvecs)
            expand = lambda x: x.expand(B, *x.shape[1:]).contiguous()  # noqa: E731
            init_v, init_j, init_o = expand(init['vertices']), expand(init['joints']), expand(init['orientations'])
        beta_ref = self._pad_ref(initial_shape_betas, B, 'initial_shape_betas')
        kid_ref = None
        if initial_kid_factor is not None and self.enable_kid:
            kid_ref = torch.as_tensor(initial_kid_factor, dtype=torch.float32, device=dev).reshape(-1).expand(B).contiguous()
        o = self._opts(num_iter, final_adjust_rots, requested_keys,
                       self._shape_weights_rule(tj, vw, jw), beta_regularizer, beta_regularizer2,
                       kid_regularizer, scale_mode, scale_regularizer)
        o.share_beta = int(bool(share_beta))
        L = _native.lib()
        s = self._struct()
        ws_bytes = L.smplfit_fit_workspace_bytes(C.byref(s), B, C.byref(o), int(tj is not None),
                                                 int(vw is not None), int(jw is not None))
        if ws_bytes == 0:
            _native.check(-2 if self._ns > 17 else -1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        p = _native.ptr
        with torch.cuda.device(dev):
            _native.check(L.smplfit_fit(
                C.byref(s), B, p(tv), p(tj), p(vw), p(jw), p(beta_ref), p(kid_ref), p(init_v), p(init_j),
                p(init_o), C.byref(o), p(rotvecs), p(out['shape_betas']), p(out['trans']),
                p(out['orientations']), p(out['relative_orientations']), p(kid), p(scale_corr), ws.data_ptr(),
                ws_bytes, _native.stream_ptr(dev),
            ))
        ws.record_stream(torch.cuda.current_stream(dev))
        if rotvecs is not None:
            out['pose_rotvecs'] = rotvecs
        if kid is not None:
            out['kid_factor'] = kid
        if scale_corr is not None:
            out['scale_corr'] = scale_corr
        return out

    # ------------------------------------------------------------------------------
    @torch.jit.unused
    def fit_from_host(self, target_vertices: torch.Tensor, target_joints: Optional[torch.Tensor] = None,
                      chunk_size: int = 512, out: Optional[dict] = None, num_iter: int = 1,
                      beta_regularizer: float = 1.0, beta_regularizer2: float = 0.0, scale_regularizer: float = 0.0,
                      kid_regularizer: Optional[float] = None, final_adjust_rots: bool = True,
                      scale_target: bool = False, scale_fit: bool = False,
                      requested_keys: Optional[list] = None) -> dict[str, torch.Tensor]:
        """``fit`` for HOST-resident float32 inputs (CPU tensors, ideally pinned) -> HOST results.

        One C-ABI call (``smplfit_fit_host``): the batch is cut into chunks of ``chunk_size`` instances, the
        host-to-device copy of chunk k+1 (library-owned copy stream) overlaps the fit of chunk k (current
        stream), and the results are copied back into pinned host tensors, so the end-to-end time approaches
        max(PCIe copy, compute) instead of their sum.  ``out`` may carry preallocated pinned result tensors
        (keys of ``fit``'s result) to reuse across calls.  The results are valid after the current stream has
        been synchronised (this method does that before returning).  Per-instance options (weights, initial
        guesses) and ``share_beta`` are not available here: use ``fit`` on device tensors."""
        if requested_keys is None:
            requested_keys = ['pose_rotvecs']
        if scale_target and scale_fit:
            raise ValueError('Only one of estim_scale_target and estim_scale_fit can be True')
        scale_mode = 1 if scale_target else (2 if scale_fit else 0)
        bm = self.body_model
        dev = bm.v_template.device
        _native.require_cuda(bm.v_template, 'the body model')
        B, V, J, S = target_vertices.shape[0], bm.num_vertices, bm.num_joints, self.n_betas

        def host(x, shape, name):
            if x is None:
                return None
            if isinstance(x, np.ndarray):
                raise TypeError(f"Expected torch.Tensor for '{name}', got numpy.ndarray.")
            if x.is_cuda:
                raise ValueError(f"'{name}' must be a CPU tensor here (use fit() for device tensors)")
            if tuple(x.shape) != tuple(shape):
                raise ValueError(f"'{name}' must have shape {tuple(shape)}, got {tuple(x.shape)}")
            return x.to(dtype=torch.float32).contiguous()

        tv = host(target_vertices, (B, V, 3), 'target_vertices')
        tj = host(target_joints, (B, J, 3), 'target_joints')
        shapes = dict(shape_betas=(B, S), trans=(B, 3), orientations=(B, J, 3, 3), relative_orientations=(B, J, 3, 3))
        if 'pose_rotvecs' in requested_keys:
            shapes['pose_rotvecs'] = (B, 3 * J)
        if self.enable_kid:
            shapes['kid_factor'] = (B,)
        if scale_mode:
            shapes['scale_corr'] = (B,)
        res = {}
        for key, shape in shapes.items():
            t = None if out is None else out.get(key)
            if t is None:
                t = torch.empty(shape, dtype=torch.float32, pin_memory=True)
            elif tuple(t.shape) != shape or t.dtype != torch.float32 or t.is_cuda or not t.is_contiguous():
                raise ValueError(f"out['{key}'] must be a contiguous float32 CPU tensor of shape {shape}")
            res[key] = t
        if B == 0:
            return res
        o = self._opts(num_iter, final_adjust_rots, requested_keys, False, beta_regularizer, beta_regularizer2,
                       kid_regularizer, scale_mode, scale_regularizer)
        L = _native.lib()
        s = self._struct()
        chunk = max(1, min(int(chunk_size), B))
        ws_bytes = L.smplfit_fit_host_workspace_bytes(C.byref(s), B, chunk, C.byref(o), int(tj is not None))
        if ws_bytes == 0:
            _native.check(-2 if self._ns > 17 else -1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        hp = lambda k: res[k].data_ptr() if k in res else 0  # noqa: E731
        with torch.cuda.device(dev):
            _native.check(L.smplfit_fit_host(
                C.byref(s), B, chunk, tv.data_ptr(), 0 if tj is None else tj.data_ptr(), C.byref(o),
                hp('pose_rotvecs'), hp('shape_betas'), hp('trans'), hp('orientations'), hp('relative_orientations'),
                hp('kid_factor'), hp('scale_corr'), ws.data_ptr(), ws_bytes, _native.stream_ptr(dev),
            ))
            torch.cuda.current_stream(dev).synchronize()  # the host tensors are the result: make them valid
        return res

    # ------------------------------------------------------------------------------
    @torch.jit.unused
    def fit_with_known_pose(
        self,
        pose_rotvecs: torch.Tensor,
        target_vertices: torch.Tensor,
        target_joints: Optional[torch.Tensor] = None,
        vertex_weights: Optional[torch.Tensor] = None,
        joint_weights: Optional[torch.Tensor] = None,
        beta_regularizer: float = 1,
        beta_regularizer2: float = 0,
        scale_regularizer: float = 0,
        kid_regularizer: Optional[float] = None,
        share_beta: bool = False,
        scale_target: bool = False,
        scale_fit: bool = False,
        beta_regularizer_reference: Optional[torch.Tensor] = None,
        kid_regularizer_reference: Optional[torch.Tensor] = None,
        requested_keys: Optional[list] = None,
    ) -> dict[str, torch.Tensor]:
        """Shape and translation for a known pose (pt/bodyfitter.py:552-653).  Differentiable like the reference's:
        with a ``requires_grad`` input the values still come from the CUDA path and the gradient from ``_adjoint``."""
        if scale_target and scale_fit:
            raise ValueError('Only one of estim_scale_target and estim_scale_fit can be True')
        tensors = [pose_rotvecs, target_vertices, target_joints, vertex_weights, joint_weights,
                   beta_regularizer_reference, kid_regularizer_reference]
        if torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad for x in tensors):
            from . import _adjoint

            opts = dict(beta_regularizer=beta_regularizer, beta_regularizer2=beta_regularizer2,
                        scale_regularizer=scale_regularizer, kid_regularizer=kid_regularizer, share_beta=share_beta,
                        scale_target=scale_target, scale_fit=scale_fit)
            names = ['pose_rotvecs', 'target_vertices', 'target_joints', 'vertex_weights', 'joint_weights',
                     'beta_regularizer_reference', 'kid_regularizer_reference']
            keys = ['shape_betas', 'trans', 'relative_orientations'] + (['kid_factor'] if self.enable_kid else []) + (
                ['scale_corr'] if (scale_target or scale_fit) else [])
            tensors = [x if (x is None or isinstance(x, torch.Tensor)) else torch.as_tensor(x) for x in tensors]
            if tensors[6] is not None:
                tensors[6] = tensors[6].reshape(-1)
            bm = self.body_model
            return _adjoint.differentiable_call(
                lambda *xs: self.fit_with_known_pose(**dict(zip(names, xs)), **opts),
                lambda *xs: _adjoint.fit_with_known_pose(bm, self.n_betas, self.enable_kid, *xs, **opts),
                keys, target_vertices.shape[0], bm.v_template.device,
                4.0 * bm.num_vertices * 3 * (self.n_betas + 14) * 2.5, bool(share_beta), tensors)
        scale_mode = 1 if scale_target else (2 if scale_fit else 0)
        bm = self.body_model
        dev = bm.v_template.device
        _native.require_cuda(bm.v_template, 'the body model')
        if target_vertices.ndim != 3:
            raise ValueError(f'target_vertices must be (batch, {bm.num_vertices}, 3)')
        B, V, J, S = target_vertices.shape[0], bm.num_vertices, bm.num_joints, self.n_betas
        tv = self._prep(target_vertices, (B, V, 3), 'target_vertices')
        tj = self._prep(target_joints, (B, J, 3), 'target_joints')
        vw = self._prep(vertex_weights, (B, V), 'vertex_weights')
        jw = self._prep(joint_weights, (B, J), 'joint_weights')
        if pose_rotvecs.shape[0] not in (1, B):
            raise ValueError(f"'pose_rotvecs' must have batch size {B} (or 1), got {tuple(pose_rotvecs.shape)}")
        new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)  # noqa: E731
        out = dict(shape_betas=new(B, S), trans=new(B, 3), relative_orientations=new(B, J, 3, 3))
        kid = new(B) if self.enable_kid else None
        scale_corr = new(B) if scale_mode else None
        if B == 0:
            if kid is not None:
                out['kid_factor'] = kid
            if scale_corr is not None:
                out['scale_corr'] = scale_corr
            return out
        glob = bm(pose_rotvecs=pose_rotvecs, return_vertices=False)['orientations']
        glob = glob.expand(B, J, 3, 3).contiguous()
        beta_ref = self._pad_ref(beta_regularizer_reference, B, 'beta_regularizer_reference')
        kid_ref = None
        if kid_regularizer_reference is not None and self.enable_kid:
            kid_ref = torch.as_tensor(kid_regularizer_reference, dtype=torch.float32, device=dev).reshape(-1)
            if kid_ref.numel() not in (1, B):
                raise ValueError(f"'kid_regularizer_reference' must have {B} (or 1) elements, got {kid_ref.numel()}")
            kid_ref = kid_ref.expand(B).contiguous()
        o = self._opts(1, False, [], self._shape_weights_rule(tj, vw, jw), beta_regularizer, beta_regularizer2,
                       kid_regularizer, scale_mode, scale_regularizer)
        o.share_beta = int(bool(share_beta))
        L = _native.lib()
        s = self._struct()
        ws_bytes = L.smplfit_fit_workspace_bytes(C.byref(s), B, C.byref(o), int(tj is not None),
                                                 int(vw is not None), int(jw is not None))
        if ws_bytes == 0:
            _native.check(-2 if self._ns > 17 else -1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        p = _native.ptr
        with torch.cuda.device(dev):
            _native.check(L.smplfit_fit_known_pose(
                C.byref(s), B, p(glob), p(tv), p(tj), p(vw), p(jw), p(beta_ref), p(kid_ref), C.byref(o),
                p(out['shape_betas']), p(out['trans']), p(out['relative_orientations']), p(kid), p(scale_corr),
                ws.data_ptr(), ws_bytes, _native.stream_ptr(dev),
            ))
        ws.record_stream(torch.cuda.current_stream(dev))
        if kid is not None:
            out['kid_factor'] = kid
        if scale_corr is not None:
            out['scale_corr'] = scale_corr
        return out

    @torch.jit.unused
    def fit_with_known_shape(
        self,
        shape_betas: torch.Tensor,
        target_vertices: torch.Tensor,
        target_joints: Optional[torch.Tensor] = None,
        vertex_weights: Optional[torch.Tensor] = None,
        joint_weights: Optional[torch.Tensor] = None,
        kid_factor: Optional[torch.Tensor] = None,
        num_iter: int = 1,
        final_adjust_rots: bool = True,
        initial_pose_rotvecs: Optional[torch.Tensor] = None,
        scale_fit: bool = False,
        requested_keys: Optional[list] = None,
    ) -> dict[str, torch.Tensor]:
        """Pose and translation for known betas (pt/bodyfitter.py:656-838).  Differentiable like the reference's (values
        from the CUDA path, gradient from ``_adjoint``)."""
        if requested_keys is None:
            requested_keys = ['pose_rotvecs']
        tensors = [shape_betas, target_vertices, target_joints, vertex_weights, joint_weights, kid_factor,
                   initial_pose_rotvecs]
        if torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad for x in tensors):
            from . import _adjoint

            opts = dict(num_iter=num_iter, final_adjust_rots=final_adjust_rots, scale_fit=scale_fit)
            names = ['shape_betas', 'target_vertices', 'target_joints', 'vertex_weights', 'joint_weights', 'kid_factor',
                     'initial_pose_rotvecs']
            want_rv = 'pose_rotvecs' in requested_keys
            want_rel = want_rv or 'relative_orientations' in requested_keys
            keys = ['trans', 'orientations'] + (['scale_corr'] if scale_fit else []) + (
                ['relative_orientations'] if want_rel else []) + (['pose_rotvecs'] if want_rv else [])
            tensors = [x if (x is None or isinstance(x, torch.Tensor)) else torch.as_tensor(x) for x in tensors]
            if tensors[5] is not None:
                tensors[5] = tensors[5].reshape(-1)
            bm = self.body_model
            return _adjoint.differentiable_call(
                lambda *xs: self.fit_with_known_shape(**dict(zip(names, xs)), **opts, requested_keys=requested_keys),
                lambda *xs: _adjoint.fit_with_known_shape(bm, self.n_betas, *xs, **opts, want_pose_rotvecs=want_rv,
                                                          want_rel_orient=want_rel),
                keys, target_vertices.shape[0], bm.v_template.device,
                4.0 * bm.num_vertices * 3 * 30 * max(1, num_iter), False, tensors)
        bm = self.body_model
        dev = bm.v_template.device
        _native.require_cuda(bm.v_template, 'the body model')
        B, V, J = target_vertices.shape[0], bm.num_vertices, bm.num_joints
        tv = self._prep(target_vertices, (B, V, 3), 'target_vertices')
        tj = self._prep(target_joints, (B, J, 3), 'target_joints')
        vw = self._prep(vertex_weights, (B, V), 'vertex_weights')
        jw = self._prep(joint_weights, (B, J), 'joint_weights')
        if kid_factor is not None and not self.enable_kid:
            raise NotImplementedError('kid_factor in fit_with_known_shape needs a fitter built with enable_kid=True')
        betas = shape_betas.to(device=dev, dtype=torch.float32)
        if betas.ndim != 2 or betas.shape[0] not in (1, B):
            raise ValueError(f"'shape_betas' must be ({B} or 1, n_betas), got {tuple(betas.shape)}")
        if B == 0:
            out = dict(trans=betas.new_empty(0, 3), orientations=betas.new_empty(0, J, 3, 3))
            if scale_fit:
                out['scale_corr'] = betas.new_empty(0)
            if 'relative_orientations' in requested_keys or 'pose_rotvecs' in requested_keys:
                out['relative_orientations'] = betas.new_empty(0, J, 3, 3)
            if 'pose_rotvecs' in requested_keys:
                out['pose_rotvecs'] = betas.new_empty(0, 3 * J)
            return out
        if betas.shape[0] != B:
            betas = betas.expand(B, betas.shape[1])
        betas = betas.contiguous()
        kid = None
        if kid_factor is not None:
            kid = torch.as_tensor(kid_factor, dtype=torch.float32, device=dev).reshape(-1).expand(B).contiguous()
        init = bm(shape_betas=betas, kid_factor=kid, pose_rotvecs=initial_pose_rotvecs)
        expand = lambda x: x.expand(B, *x.shape[1:]).contiguous()  # noqa: E731
        init_v, init_j, init_o = expand(init['vertices']), expand(init['joints']), expand(init['orientations'])
        new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)  # noqa: E731
        out = dict(trans=new(B, 3), orientations=new(B, J, 3, 3))
        want_rel = 'relative_orientations' in requested_keys or 'pose_rotvecs' in requested_keys
        rel = new(B, J, 3, 3) if want_rel else None
        rotvecs = new(B, 3 * J) if 'pose_rotvecs' in requested_keys else None
        scale_corr = new(B) if scale_fit else None
        o = self._opts(num_iter, final_adjust_rots, requested_keys, False, 0.0, 0.0, None, 2 if scale_fit else 0, 0.0)
        L = _native.lib()
        s = self._struct()
        ws_bytes = L.smplfit_fit_workspace_bytes(C.byref(s), B, C.byref(o), int(tj is not None),
                                                 int(vw is not None), int(jw is not None))
        if ws_bytes == 0:
            _native.check(-2 if self._ns > 17 else -1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        p = _native.ptr
        with torch.cuda.device(dev):
            _native.check(L.smplfit_fit_known_shape(
                C.byref(s), B, p(betas), betas.shape[1], p(kid), p(tv), p(tj), p(vw), p(jw), p(init_v), p(init_j),
                p(init_o), C.byref(o), p(rotvecs), p(out['trans']), p(out['orientations']), p(rel), p(scale_corr),
                ws.data_ptr(), ws_bytes, _native.stream_ptr(dev),
            ))
        ws.record_stream(torch.cuda.current_stream(dev))
        if scale_corr is not None:
            out['scale_corr'] = scale_corr
        if rel is not None:
            out['relative_orientations'] = rel
        if rotvecs is not None:
            out['pose_rotvecs'] = rotvecs
        return out
