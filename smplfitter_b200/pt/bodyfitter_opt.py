"""``BodyFitterOpt``: the closed-form fit followed by an optional first-order refinement
(/root/reference/src/smplfitter/pt/bodyfitter_opt.py:35-255; same constructor, ``fit`` arguments and result keys).

The refinement minimises the mean vertex (+ joint) distance with Adam over the GLOBAL joint rotations in the
continuous 6-D parametrisation (two columns, Gram-Schmidt), the betas and the translation.  Every step is one CUDA
forward pass (``BodyModel.forward`` with ``glob_rotmats``) and its registered backward (pt/_adjoint.py).
"""

from __future__ import annotations

import contextlib
import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from .bodyfitter import BodyFitter
from .rotation import mat2rotvec


def rot6d_to_rotmat(x: torch.Tensor) -> torch.Tensor:
    """(..., 6) -> (..., 3, 3): columns = Gram-Schmidt of the two stored columns and their cross product
    (bodyfitter_opt.py:16-27, same 1e-8 guard on the norms)."""
    u, v = x[..., :3], x[..., 3:6]
    e1 = u / (u.norm(dim=-1, keepdim=True) + 1e-8)
    v = v - (e1 * v).sum(-1, keepdim=True) * e1
    e2 = v / (v.norm(dim=-1, keepdim=True) + 1e-8)
    return torch.stack([e1, e2, torch.linalg.cross(e1, e2)], dim=-1)


def rotmat_to_rot6d(R: torch.Tensor) -> torch.Tensor:
    """First two columns (bodyfitter_opt.py:30-32)."""
    return torch.cat([R[..., :, 0], R[..., :, 1]], dim=-1)


def _learning_rate(step: int, num_steps: int, lr: float, warmup_ratio: float) -> float:
    """Linear warm-up, then half a cosine down to zero (bodyfitter_opt.py:181-186)."""
    warm = int(num_steps * warmup_ratio)
    if step < warm:
        return lr * (step + 1) / warm
    return lr * 0.5 * (1.0 + math.cos(math.pi * (step - warm) / max(1, num_steps - warm)))


def _mean_distance(x: torch.Tensor, target: torch.Tensor, w: Optional[torch.Tensor]) -> torch.Tensor:
    d = torch.linalg.norm(x - target, dim=-1)
    return d.mean() if w is None else (w * d).mean()


class BodyFitterOpt(nn.Module):
    def __init__(self, body_model, enable_kid: bool = False):
        super().__init__()
        self.body_model = body_model
        self.fitter = BodyFitter(body_model, enable_kid=enable_kid)
        self.enable_kid = enable_kid

    def fit(self, target_vertices: torch.Tensor, target_joints: Optional[torch.Tensor] = None,
            vertex_weights: Optional[torch.Tensor] = None, joint_weights: Optional[torch.Tensor] = None,
            num_iter: int = 1, beta_regularizer: float = 1, beta_regularizer2: float = 0, share_beta: bool = False,
            final_adjust_rots: bool = True, scale_target: bool = False, scale_fit: bool = False,
            refine_steps: int = 0, refine_lr: float = 0.03, warmup_ratio: float = 0.5) -> Dict[str, torch.Tensor]:
        """``refine_steps == 0``: exactly ``BodyFitter.fit``.  Otherwise the closed-form result (without the final
        rotation adjustment, as in the reference) starts ``refine_steps`` Adam steps."""
        with torch.no_grad() if refine_steps else contextlib.nullcontext():
            init = self.fitter.fit(
                target_vertices, target_joints=target_joints, vertex_weights=vertex_weights, joint_weights=joint_weights,
                num_iter=num_iter, beta_regularizer=beta_regularizer, beta_regularizer2=beta_regularizer2,
                share_beta=share_beta, final_adjust_rots=final_adjust_rots if refine_steps == 0 else False,
                scale_target=scale_target, scale_fit=scale_fit, requested_keys=['pose_rotvecs', 'shape_betas', 'trans'])
        if refine_steps == 0:
            return init
        bm = self.body_model
        dev = bm.v_template.device
        to = lambda x: None if x is None else x.detach().to(dev, torch.float32)  # noqa: E731
        tv, tj, vw, jw = to(target_vertices), to(target_joints), to(vertex_weights), to(joint_weights)
        with torch.no_grad():
            glob0 = bm(pose_rotvecs=init['pose_rotvecs'], return_vertices=False)['orientations']
        rot6d = rotmat_to_rot6d(glob0).clone().requires_grad_(True)
        betas = init['shape_betas'].detach().clone().requires_grad_(True)
        trans = init['trans'].detach().clone().requires_grad_(True)
        kid = init['kid_factor'].detach().clone().requires_grad_(True) if 'kid_factor' in init else None
        params = [rot6d, betas, trans] + ([kid] if kid is not None else [])
        optimizer = torch.optim.Adam(params, lr=refine_lr, betas=(0.97, 0.999))
        for step in range(refine_steps):
            for group in optimizer.param_groups:
                group['lr'] = _learning_rate(step, refine_steps, refine_lr, warmup_ratio)
            optimizer.zero_grad()
            out = bm(glob_rotmats=rot6d_to_rotmat(rot6d), shape_betas=betas, trans=trans, kid_factor=kid)
            loss = _mean_distance(out['vertices'], tv, vw)
            if tj is not None:
                loss = loss + _mean_distance(out['joints'], tj, jw)
            if beta_regularizer > 0 and betas.shape[1] > 2:
                loss = loss + beta_regularizer * betas[:, 2:].pow(2).mean()
            loss.backward()
            optimizer.step()
        with torch.no_grad():
            glob = rot6d_to_rotmat(rot6d)
            par = bm.kintree_parents_tensor[1:].to(dev)
            rel = torch.cat([glob[:, :1], glob[:, par].transpose(-1, -2) @ glob[:, 1:]], dim=1)
            result = {'pose_rotvecs': mat2rotvec(rel).reshape(glob.shape[0], -1), 'shape_betas': betas.detach(),
                      'trans': trans.detach()}
        if kid is not None:
            result['kid_factor'] = kid.detach()
        return result
