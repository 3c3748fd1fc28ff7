"""``BodyModel`` -- drop-in for ``smplfitter.pt.BodyModel``
(/root/reference/src/smplfitter/pt/bodymodel.py:12-453) whose ``forward`` runs in the sm_100a
CUDA library (``smplfit_forward``).  Same constructor and method signatures, same buffers,
same result dictionaries, same error conventions (:164-200, :210-217).
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _native, masks, modeldata
from . import _ops

SEG_LEN = 32  # vertices per statistics segment (== the padded slot block of one segment in csrc/fit_fused.cu)
CHUNK_LEN = 128  # vertices per shape-pass chunk


N_SLOTS = 12  # distinct skinning joints a segment may touch (per-warp accumulator slots of k_shape_lite)


def _segments(part_sorted: np.ndarray, num_joints: int, joint_sets=None):
    """Split the part-grouped vertex order into segments of <= SEG_LEN vertices of one part that
    touch at most N_SLOTS distinct skinning joints (``joint_sets[i]`` = joints of sorted vertex i)."""
    seg_start, seg_part = [0], []
    part_seg_begin = np.zeros(num_joints + 1, dtype=np.int32)
    V = len(part_sorted)
    i = 0
    for p in range(num_joints):
        part_seg_begin[p] = len(seg_part)
        n = int((part_sorted == p).sum())
        done = 0
        while done < n:
            step = min(SEG_LEN, n - done)
            if joint_sets is not None:
                seen = set()
                for q in range(step):
                    nxt = seen | set(joint_sets[i + done + q])
                    if len(nxt) > N_SLOTS:
                        step = max(q, 1)
                        break
                    seen = nxt
            seg_part.append(p)
            seg_start.append(i + done + step)
            done += step
        i += n
    part_seg_begin[num_joints] = len(seg_part)
    assert i == V
    return np.array(seg_start, np.int32), np.array(seg_part, np.int32), part_seg_begin


def _slot_tables(seg_start, joint_sets, num_joints):
    """Per segment the sorted list of joints it touches, and per joint the (segment, slot) cells."""
    n_seg = len(seg_start) - 1
    seg_slots = np.full((n_seg, N_SLOTS), -1, np.int32)
    cells = [[] for _ in range(num_joints)]
    for sgm in range(n_seg):
        js = sorted(set().union(*[set(joint_sets[i]) for i in range(seg_start[sgm], seg_start[sgm + 1])]))
        assert len(js) <= N_SLOTS
        for slot, j in enumerate(js):
            seg_slots[sgm, slot] = j
            cells[j].append(sgm * N_SLOTS + slot)
    yj_start = np.zeros(num_joints + 1, np.int32)
    yj_start[1:] = np.cumsum([len(c) for c in cells])
    yj_entry = np.array([e for c in cells for e in c] or [0], np.int32)
    return seg_slots, yj_start, yj_entry


class BodyModel(_ops.RegisteredModule, nn.Module):
    """Statistical body model of the SMPL family (forward linear blend skinning).

    Parameters are those of the reference (pt/bodymodel.py:53-64).  Model data comes from
    ``smplfitter_b200.modeldata.initialize`` (same signature as the reference seam
    common.py:219): an installed provider for the licensed files, or the synthetic stand-in.
    """

    def __init__(
        self,
        model_name: str = 'smpl',
        gender: str = 'neutral',
        model_root: Optional[str] = None,
        num_betas: Optional[int] = None,
        vertex_subset_size: Optional[int] = None,
        vertex_subset=None,
        faces=None,
        joint_regressor_post_lbs=None,
        device=None,
    ):
        super().__init__()
        self.gender = gender
        self.model_name = model_name
        data = modeldata.initialize(
            model_name, gender, model_root, num_betas, vertex_subset_size, vertex_subset, faces,
            joint_regressor_post_lbs,
        )
        f32 = lambda x: torch.from_numpy(np.array(x, dtype=np.float32, order='C'))  # (a private copy)  # noqa: E731
        self.v_template = nn.Buffer(f32(data.v_template))
        self.shapedirs = nn.Buffer(f32(data.shapedirs))
        self.posedirs = nn.Buffer(f32(data.posedirs))
        self.J_regressor_post_lbs = nn.Buffer(f32(data.J_regressor_post_lbs))
        self.J_template = nn.Buffer(f32(data.J_template))
        self.J_shapedirs = nn.Buffer(f32(data.J_shapedirs))
        self.kid_shapedir = nn.Buffer(f32(data.kid_shapedir))
        self.kid_J_shapedir = nn.Buffer(f32(data.kid_J_shapedir))
        self.weights = nn.Buffer(f32(data.weights))
        self.kintree_parents_tensor = nn.Buffer(torch.tensor(data.kintree_parents, dtype=torch.int64))
        self.kintree_parents = data.kintree_parents
        self.faces = data.faces
        self.num_joints = data.num_joints
        self.num_vertices = data.num_vertices
        self.num_betas = self.shapedirs.shape[2]
        self.vertex_subset = data.vertex_subset
        self.joint_names = data.joint_names
        if self.vertex_subset is None:
            self.vertex_subset = np.arange(self.num_vertices)
        for i in range(1, self.num_joints):
            if not 0 <= int(self.kintree_parents[i]) < i:
                raise ValueError('kinematic tree must list parents before children')

        # ---- derived device tables (host precompute, see masks.py) ----
        w32 = self.weights.numpy()
        self._plan = masks.build_fit_plan(w32, self.kintree_parents, model_name)
        plan = self._plan
        V, J = self.num_vertices, self.num_joints
        P = 9 * (J - 1)
        skin_idx, skin_w, K = masks.sparse_skin_table(w32)
        if K <= 4:
            # internal vertex order: by part, then by the tuple of skinning joints (weight-descending slots), so that
            # consecutive vertices share their joints and the vertex kernels keep the joint rows in registers
            key_idx = np.zeros((V, 4), np.int32)
            key_w = np.zeros((V, 4), np.float32)
            key_idx[:, :K], key_w[:, :K] = skin_idx, skin_w
            key_idx[:, K:] = skin_idx[:, :1]
            key_idx = np.take_along_axis(key_idx, np.argsort(-key_w, axis=1, kind='stable'), axis=1)
            order = np.lexsort((np.arange(V), key_idx[:, 3], key_idx[:, 2], key_idx[:, 1], key_idx[:, 0],
                                plan.part_assignment)).astype(np.int32)
        else:
            order = np.argsort(plan.part_assignment, kind='stable').astype(np.int32)
        inv_order = np.empty(V, np.int32)
        inv_order[order] = np.arange(V, dtype=np.int32)
        joint_sets = [tuple(int(j) for j, w in zip(skin_idx[v], skin_w[v]) if w != 0) for v in order]
        seg_start, seg_part, part_seg_begin = _segments(plan.part_assignment[order], J, joint_sets if K <= 4 else None)
        Kp = (P + 15) // 16 * 16
        posedirs_fit = np.zeros((V * 3, Kp), np.float32)
        posedirs_fit[:, :P] = self.posedirs.numpy()[order].reshape(V * 3, P)
        Kt = (P + 31) // 32 * 32
        pd = np.zeros((V * 3, Kt), np.float32)
        pd[:, :P] = posedirs_fit[:, :P]
        pd_hi = (pd.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)  # tf32-exact part
        pd_lo = pd - pd_hi
        # the same rows for the fp16-split tensor-core GEMM (csrc/fwd_fused.cu, MODE 1): scaled into fp16's normal
        # range, hi = fp16(x), lo = fp16(x - hi); rows padded to the 192-row tile
        kf = (P + 31) // 32 * 32
        pf = np.zeros(((V * 3 + 191) // 192 * 192, kf), np.float64)
        pf[:V * 3, :P] = posedirs_fit[:, :P]
        amax = float(np.abs(pf).max())
        self._fit_scale_log2 = int(np.clip(np.floor(np.log2(16384.0 / amax)) if amax > 0 else 0, 0, 14))
        pf *= 2.0 ** self._fit_scale_log2
        pf_hi = pf.astype(np.float16)
        pf_lo = (pf - pf_hi.astype(np.float64)).astype(np.float16)
        self._fit_kf = kf
        eye_feat = np.tile(np.eye(3, dtype=np.float32), [J - 1, 1]).reshape(-1)
        v_posed0 = self.v_template.numpy() + np.einsum('vcp,p->vc', self.posedirs.numpy(), eye_feat)
        template_mesh = (v_posed0 * w32.sum(axis=1, keepdims=True)).astype(np.float32)  # pt/bodyfitter.py:49
        flags = (plan.part_is_stat.astype(np.int32) | (plan.part_is_adjustable.astype(np.int32) << 1))
        jreg = self.J_regressor_post_lbs.numpy()
        i32 = lambda x: torch.from_numpy(np.array(x, dtype=np.int32, order='C'))  # noqa: E731
        t = {
            'parents': i32(plan.parents), 'skin_idx': i32(skin_idx), 'skin_w': f32(skin_w),
            'order': i32(order), 'inv_order': i32(inv_order), 'seg_start': i32(seg_start),
            'seg_part': i32(seg_part), 'part_seg_begin': i32(part_seg_begin),
            'part_kind': i32(plan.part_kind), 'part_copy_src': i32(plan.part_copy_src),
            'part_flags': i32(flags), 'cas_table': i32(plan.cas_table), 'cas_count': i32(plan.cas_count),
            'posedirs_fit': f32(posedirs_fit), 'v_template_fit': f32(self.v_template.numpy()[order].reshape(-1)),
            'template_mesh': f32(template_mesh),
            'template_joints_regressed': f32(jreg @ template_mesh),
            'J_regressor_fit': f32(jreg[:, order]),
            'posedirs_hi': f32(pd_hi), 'posedirs_lo': f32(pd_lo),
            'fit_P_hi': torch.from_numpy(pf_hi), 'fit_P_lo': torch.from_numpy(pf_lo),
            'template_mesh_fit': f32(template_mesh[order]),
        }
        jr = jreg[:, order]
        nzr, nzc = np.nonzero(jr)  # row-major: grouped by joint, ascending internal position
        jptr = np.zeros(J + 1, np.int32)
        jptr[1:] = np.cumsum(np.bincount(nzr, minlength=J))
        t['jreg_ptr'], t['jreg_idx'] = i32(jptr), i32(nzc if len(nzc) else np.zeros(1, np.int32))
        t['jreg_val'] = f32(jr[nzr, nzc] if len(nzc) else np.zeros(1, np.float32))
        if K <= 4 and J <= 64:
            v_rest = self.v_template.numpy().astype(np.float64) + np.einsum(
                'vcp,p->vc', self.posedirs.numpy().astype(np.float64), eye_feat.astype(np.float64))
            self._build_forward_tables(skin_idx, skin_w, v_rest, t)
        if K <= 4:
            seg_slots, yj_start, yj_entry = _slot_tables(seg_start, joint_sets, J)
            t['seg_slots'], t['yj_start'], t['yj_entry'] = i32(seg_slots), i32(yj_start), i32(yj_entry)
        for k, v in t.items():
            self.register_buffer('_t_' + k, v, persistent=False)
        self._dims = dict(
            num_vertices=V, num_joints=J, num_betas=self.num_betas, num_pose_feats=P, skin_k=K,
            is_smpl_family=int(plan.is_smpl_family), n_used=int(plan.part_is_stat[plan.part_assignment].sum()),
            n_segments=len(seg_part), chunk_len=CHUNK_LEN, max_cas=plan.cas_table.shape[1],
            n_adjustable=int(np.asarray(plan.part_is_adjustable).astype(bool).sum()),
        )
        self._handle = _ops.register(self)
        if device is not None:
            self.to(device)

    @torch.jit.unused
    def _build_forward_tables(self, skin_idx, skin_w, v_rest, t):
        """Constants of the fused forward kernel (include/smplfit_b200.h ``fwd_*``; csrc/fwd_fused.cu).

        Processing order: 64-vertex tiles of 16-vertex chunks; a chunk is 16 consecutive model vertices (so a chunk's
        results are one contiguous 192-byte piece of the caller's row), visited in the order and with the joint -> slot
        assignment that makes the kernel's four register-cached joint rows change as rarely as possible (greedy: next
        the vertex needing the fewest new joints; a new joint evicts the slot whose joint is needed latest)."""
        V, J, S = self.num_vertices, self.num_joints, self.num_betas
        P = 9 * (J - 1)
        CH, TILE = 16, 64  # FWD_CHUNK and TILE_V of csrc/fwd_fused.cu
        Vp = (V + TILE - 1) // TILE * TILE
        joints_of = [[(int(j), float(w)) for j, w in zip(skin_idx[v], skin_w[v]) if w != 0] for v in range(V)]
        rec = np.zeros((Vp, 8), np.uint32)
        proc = np.full(Vp, -1, np.int64)  # processing position -> model vertex (-1: padding)
        # one sequential replay of the cache over the whole processing order: every record carries the cache content
        # AFTER its vertex (the joint of each slot), so a warp entering the order anywhere loads all four slots from the
        # record of its first vertex (the kernel does that at the start of its share of every tile)
        v_rest32 = v_rest.astype(np.float32).view(np.uint32)
        cache = [-1, -1, -1, -1]
        for c0 in range(0, Vp, CH):
            todo = [v for v in range(c0, min(c0 + CH, V))]
            pos = c0
            while todo:
                # fewest joints missing from the cache; ties: lowest vertex index
                v = min(todo, key=lambda u: (sum(1 for j, _ in joints_of[u] if j not in cache), u))
                todo.remove(v)
                need = [j for j, _ in joints_of[v]]
                reload = 0
                for j in need:
                    if j in cache:
                        continue
                    free = [k for k in range(4) if cache[k] not in need]
                    # evict an empty slot, else the slot whose joint the fewest remaining vertices of the chunk need
                    k = min(free, key=lambda kk: (cache[kk] != -1,
                                                  sum(1 for u in todo for jj, _ in joints_of[u] if jj == cache[kk]), kk))
                    cache[k] = j
                    reload |= 1 << k
                w = np.zeros(4, np.float32)
                for j, ww in joints_of[v]:
                    w[cache.index(j)] = ww
                pack = 0
                for k in range(4):
                    pack |= (max(cache[k], 0) & 63) << (6 * k)
                pack |= reload << 24
                pack |= (v - c0) << 28
                rec[pos, 0:4] = w.view(np.uint32)
                rec[pos, 4] = pack
                rec[pos, 5:8] = v_rest32[v]
                proc[pos] = v
                pos += 1
            for q in range(pos, c0 + CH):  # padding of the last chunk(s): zero weights, distinct local indices
                pack = 0
                for k in range(4):
                    pack |= (max(cache[k], 0) & 63) << (6 * k)
                rec[q, 4] = pack | ((q - c0) << 28)
        Kf = (P + S + 1 + 31) // 32 * 32
        pm = np.zeros((Vp * 3, Kf), np.float64)
        live = proc >= 0
        rows = np.concatenate([self.posedirs.numpy().astype(np.float64), self.shapedirs.numpy().astype(np.float64),
                               self.kid_shapedir.numpy().astype(np.float64)[:, :, None]], axis=2)  # (V,3,P+S+1)
        pm.reshape(Vp, 3, Kf)[live, :, :P + S + 1] = rows[proc[live]]
        amax = float(np.abs(pm).max())
        self._fwd_scale_log2 = int(np.clip(np.floor(np.log2(16384.0 / amax)) if amax > 0 else 0, 0, 14))
        pm *= 2.0 ** self._fwd_scale_log2
        hi = pm.astype(np.float16)
        lo = (pm - hi.astype(np.float64)).astype(np.float16)
        self._fwd_kf = Kf
        t['fwd_P_hi'] = torch.from_numpy(hi)
        t['fwd_P_lo'] = torch.from_numpy(lo)
        t['fwd_vrec'] = torch.from_numpy(rec.view(np.int32))

    # ------------------------------------------------------------------------------
    @torch.jit.unused
    def _struct(self, extra: Optional[dict] = None):
        """Fill the C struct with the current device addresses of the buffers."""
        _native.require_cuda(self.v_template, 'the body model')
        s = _native.ModelStruct()
        for k, v in self._dims.items():
            setattr(s, k, v)
        s.v_template = self.v_template.data_ptr()
        s.shapedirs = self.shapedirs.data_ptr()
        s.posedirs = self.posedirs.data_ptr()
        s.kid_shapedir = self.kid_shapedir.data_ptr()
        s.J_template = self.J_template.data_ptr()
        s.J_shapedirs = self.J_shapedirs.data_ptr()
        s.kid_J_shapedir = self.kid_J_shapedir.data_ptr()
        s.J_regressor = self.J_regressor_post_lbs.data_ptr()
        for name in ('parents', 'skin_idx', 'skin_w', 'order', 'inv_order', 'seg_start', 'seg_part',
                     'part_seg_begin', 'part_kind', 'part_copy_src', 'part_flags', 'cas_table', 'cas_count',
                     'posedirs_fit', 'v_template_fit', 'template_mesh', 'template_joints_regressed',
                     'J_regressor_fit', 'posedirs_hi', 'posedirs_lo', 'template_mesh_fit', 'fit_P_hi', 'fit_P_lo',
                     'jreg_ptr', 'jreg_idx', 'jreg_val'):
            setattr(s, name, getattr(self, '_t_' + name).data_ptr())
        for name, buf in self.named_buffers():
            if not buf.is_contiguous():
                raise RuntimeError(f'smplfitter_b200: buffer {name} must be contiguous')
        s.fit_ns = 0
        s.fit_kf, s.fit_scale_log2 = self._fit_kf, self._fit_scale_log2
        if hasattr(self, '_t_fwd_P_hi'):
            s.fwd_P_hi, s.fwd_P_lo, s.fwd_vrec = (self._t_fwd_P_hi.data_ptr(), self._t_fwd_P_lo.data_ptr(),
                                                  self._t_fwd_vrec.data_ptr())
            s.fwd_kf, s.fwd_scale_log2 = self._fwd_kf, self._fwd_scale_log2
        if hasattr(self, '_t_seg_slots'):
            s.seg_slots, s.yj_start, s.yj_entry = (self._t_seg_slots.data_ptr(), self._t_yj_start.data_ptr(),
                                                    self._t_yj_entry.data_ptr())
            s.n_slots = N_SLOTS
        if extra:
            for k, v in extra.items():
                setattr(s, k, v)
        return s

    # ------------------------------------------------------------------------------
    def forward(
        self,
        pose_rotvecs: Optional[torch.Tensor] = None,
        shape_betas: Optional[torch.Tensor] = None,
        trans: Optional[torch.Tensor] = None,
        kid_factor: Optional[torch.Tensor] = None,
        rel_rotmats: Optional[torch.Tensor] = None,
        glob_rotmats: Optional[torch.Tensor] = None,
        return_vertices: bool = True,
    ) -> Dict[str, torch.Tensor]:
        """Vertices, joints and global orientations for a batch (pt/bodymodel.py:121-307).

        TorchScript-compatible: dispatches through the ``smplfit_b200::forward`` custom op."""
        if not torch.jit.is_scripting():
            # eager callers get the reference's TypeError / ValueError before the op schema sees the arguments
            self._check_forward_inputs(pose_rotvecs, shape_betas, trans, kid_factor, rel_rotmats, glob_rotmats)
        outs = torch.ops.smplfit_b200.forward(self._handle, self.v_template, pose_rotvecs, shape_betas, trans,
                                              kid_factor, rel_rotmats, glob_rotmats, return_vertices)
        result: Dict[str, torch.Tensor] = {'joints': outs[0], 'orientations': outs[1]}
        if return_vertices:
            result['vertices'] = outs[2]
        return result

    @torch.jit.unused
    def _check_forward_inputs(self, pose_rotvecs, shape_betas, trans, kid_factor, rel_rotmats, glob_rotmats):
        """Argument checks of pt/bodymodel.py:173-200 (skipped under TorchScript, as in the reference)."""
        n_rot = sum(x is not None for x in (pose_rotvecs, rel_rotmats, glob_rotmats))
        if n_rot > 1:
            raise ValueError(
                'Only one rotation input may be provided (pose_rotvecs, rel_rotmats, or glob_rotmats).'
            )
        for name, arg, min_ndim in [
            ('pose_rotvecs', pose_rotvecs, 2), ('shape_betas', shape_betas, 2), ('trans', trans, 2),
            ('kid_factor', kid_factor, 1), ('rel_rotmats', rel_rotmats, 4), ('glob_rotmats', glob_rotmats, 4),
        ]:
            if arg is not None:
                if isinstance(arg, np.ndarray):
                    raise TypeError(
                        f"Expected torch.Tensor for '{name}', got numpy.ndarray. "
                        f'Convert with torch.from_numpy() or torch.as_tensor().'
                    )
                if arg.ndim < min_ndim:
                    raise ValueError(
                        f"Expected batched input for '{name}' with at least {min_ndim} dimensions, but "
                        f'got shape {tuple(arg.shape)}. For single (unbatched) inputs, use model.single() '
                        f'instead.'
                    )

    @torch.jit.unused
    def _forward_impl(
        self,
        pose_rotvecs: Optional[torch.Tensor] = None,
        shape_betas: Optional[torch.Tensor] = None,
        trans: Optional[torch.Tensor] = None,
        kid_factor: Optional[torch.Tensor] = None,
        rel_rotmats: Optional[torch.Tensor] = None,
        glob_rotmats: Optional[torch.Tensor] = None,
        return_vertices: bool = True,
    ) -> Dict[str, torch.Tensor]:
        n_rot = sum(x is not None for x in (pose_rotvecs, rel_rotmats, glob_rotmats))
        if n_rot > 1:
            raise ValueError(
                'Only one rotation input may be provided (pose_rotvecs, rel_rotmats, or glob_rotmats).'
            )
        for name, arg, min_ndim in [
            ('pose_rotvecs', pose_rotvecs, 2), ('shape_betas', shape_betas, 2), ('trans', trans, 2),
            ('kid_factor', kid_factor, 1), ('rel_rotmats', rel_rotmats, 4), ('glob_rotmats', glob_rotmats, 4),
        ]:
            if arg is not None:
                if isinstance(arg, np.ndarray):
                    raise TypeError(
                        f"Expected torch.Tensor for '{name}', got numpy.ndarray. "
                        f'Convert with torch.from_numpy() or torch.as_tensor().'
                    )
                if arg.ndim < min_ndim:
                    raise ValueError(
                        f"Expected batched input for '{name}' with at least {min_ndim} dimensions, but "
                        f'got shape {tuple(arg.shape)}. For single (unbatched) inputs, use model.single() '
                        f'instead.'
                    )
        batch_size = 0
        for arg in (pose_rotvecs, shape_betas, trans, rel_rotmats, glob_rotmats):
            if arg is not None:
                batch_size = arg.shape[0]
                break
        device = self.v_template.device
        J, V = self.num_joints, self.num_vertices
        if batch_size == 0:
            result = dict(
                joints=torch.empty((0, J, 3), device=device),
                orientations=torch.empty((0, J, 3, 3), device=device),
            )
            if return_vertices:
                result['vertices'] = torch.empty((0, V, 3), device=device)
            return result
        _native.require_cuda(self.v_template, 'the body model')

        def prep(x, shape=None):
            x = x.to(device=device, dtype=torch.float32)
            if shape is not None:
                x = x.reshape(shape)
            return x.contiguous()

        if rel_rotmats is not None:
            rot, mode = prep(rel_rotmats, (batch_size, J, 3, 3)), 1
        elif pose_rotvecs is not None:
            rot, mode = prep(pose_rotvecs, (batch_size, J * 3)), 0
        elif glob_rotmats is not None:
            rot, mode = prep(glob_rotmats, (batch_size, J, 3, 3)), 2
        else:
            rot, mode = None, 3
        betas = prep(shape_betas) if shape_betas is not None else None
        n_betas = 0 if betas is None else betas.shape[1]
        if betas is not None and betas.shape[0] != batch_size:
            betas = betas.expand(batch_size, n_betas).contiguous()
        tr = prep(trans) if trans is not None else None
        if tr is not None and tr.shape[0] != batch_size:
            tr = tr.expand(batch_size, 3).contiguous()
        kid = None
        if kid_factor is not None:
            kid = torch.as_tensor(kid_factor, dtype=torch.float32, device=device).reshape(-1)
            kid = kid.expand(batch_size).contiguous()
        joints = torch.empty((batch_size, J, 3), device=device, dtype=torch.float32)
        orient = torch.empty((batch_size, J, 3, 3), device=device, dtype=torch.float32)
        verts = torch.empty((batch_size, V, 3), device=device, dtype=torch.float32) if return_vertices else None
        L = _native.lib()
        s = self._struct()
        ws_bytes = L.smplfit_forward_workspace_bytes(C.byref(s), batch_size)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            _native.check(L.smplfit_forward(
                C.byref(s), batch_size, mode, _native.ptr(rot), _native.ptr(betas), n_betas, _native.ptr(tr),
                _native.ptr(kid), _native.ptr(verts), _native.ptr(joints), _native.ptr(orient),
                ws.data_ptr(), ws_bytes, _native.stream_ptr(device),
            ))
        ws.record_stream(torch.cuda.current_stream(device))
        result = dict(joints=joints, orientations=orient)
        if return_vertices:
            result['vertices'] = verts
        return result

    @torch.jit.unused
    def single(
        self,
        pose_rotvecs: Optional[torch.Tensor] = None,
        shape_betas: Optional[torch.Tensor] = None,
        trans: Optional[torch.Tensor] = None,
        kid_factor: Optional[torch.Tensor] = None,
        rel_rotmats: Optional[torch.Tensor] = None,
        glob_rotmats: Optional[torch.Tensor] = None,
        return_vertices: bool = True,
    ) -> dict[str, torch.Tensor]:
        """Unbatched variant (pt/bodymodel.py:309-380)."""
        un = lambda x: x.unsqueeze(0) if x is not None else None  # noqa: E731
        pose_rotvecs, shape_betas, trans = un(pose_rotvecs), un(shape_betas), un(trans)
        rel_rotmats, glob_rotmats = un(rel_rotmats), un(glob_rotmats)
        if all(x is None for x in (pose_rotvecs, shape_betas, trans, rel_rotmats, glob_rotmats)):
            shape_betas = torch.zeros((1, 0), dtype=torch.float32, device=self.v_template.device)
        result = self.forward(
            pose_rotvecs=pose_rotvecs, shape_betas=shape_betas, trans=trans, kid_factor=kid_factor,
            rel_rotmats=rel_rotmats, glob_rotmats=glob_rotmats, return_vertices=return_vertices,
        )
        return {k: v.squeeze(0) for k, v in result.items()}

    @torch.jit.unused
    def rototranslate(
        self,
        R: torch.Tensor,
        t: Optional[torch.Tensor] = None,
        pose_rotvecs: Optional[torch.Tensor] = None,
        shape_betas: Optional[torch.Tensor] = None,
        trans: Optional[torch.Tensor] = None,
        kid_factor: Optional[torch.Tensor] = None,
        post_translate: bool = True,
    ) -> tuple[torch.Tensor, torch.Tensor]:
        """Rotate/translate a body in parametric form (pt/bodymodel.py:382-453).

        A handful of 3-vector operations on one instance: host-side tensor algebra, not part of
        the batched hot path.
        """
        from .rotation import mat2rotvec, rotvec2mat

        missing = [n for n, v in (('pose_rotvecs', pose_rotvecs), ('shape_betas', shape_betas), ('trans', trans)) if v is None]
        if missing:
            raise ValueError('pose_rotvecs, shape_betas, and trans are required.')
        shift = R.new_zeros(3) if t is None else t
        # the root joint turns with R about the rest pelvis p(beta): x -> R (x - p) + p, so the body translation
        # picks up (R - I) p on top of the rigid motion of `trans`
        n_b = shape_betas.shape[0]
        p = self.J_template[0] + self.J_shapedirs[0, :, :n_b] @ shape_betas
        if kid_factor is not None:
            p = p + kid_factor * self.kid_J_shapedir[0]
        root = mat2rotvec(R @ rotvec2mat(pose_rotvecs[:3]))
        moved = R @ trans + shift if post_translate else R @ (trans - shift)
        return torch.cat([root, pose_rotvecs[3:]]), moved + (R @ p - p)
