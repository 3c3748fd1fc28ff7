"""Single-instance rotation helpers for the non-batched convenience methods
(``BodyModel.rototranslate``).  The batched hot path does this in CUDA (csrc/linalg.cuh);
conventions follow /root/reference/src/smplfitter/pt/rotation.py:236-289.
"""

from __future__ import annotations

import torch


def rotvec2mat(rotvec: torch.Tensor) -> torch.Tensor:
    """Rodrigues formula, (..., 3) -> (..., 3, 3); zero vector -> identity."""
    angle = torch.linalg.norm(rotvec, dim=-1, keepdim=True)
    axis = torch.where(angle == 0, torch.zeros_like(rotvec), rotvec / torch.where(angle == 0, torch.ones_like(angle), angle))
    x, y, z = axis.unbind(-1)
    zero = torch.zeros_like(x)
    K = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], dim=-1).unflatten(-1, (3, 3))
    s = torch.sin(angle).unsqueeze(-1)
    c = torch.cos(angle).unsqueeze(-1)
    eye = torch.eye(3, dtype=rotvec.dtype, device=rotvec.device)
    outer = axis.unsqueeze(-1) * axis.unsqueeze(-2)
    return c * eye + s * K + (1 - c) * outer


def mat2rotvec(R: torch.Tensor) -> torch.Tensor:
    """Log map through the quaternion; branch order of pt/rotation.py:280-285."""
    r = R.flatten(-2, -1)
    r00, r01, r02, r10, r11, r12, r20, r21, r22 = r.unbind(-1)
    trace = r00 + r11 + r22
    q0 = torch.stack([r21 - r12, r02 - r20, r10 - r01, 1 + trace], -1)
    q1 = torch.stack([(1 - r22) + (r00 - r11), r10 + r01, r02 + r20, r21 - r12], -1)
    q2 = torch.stack([r10 + r01, (1 - r22) - (r00 - r11), r21 + r12, r02 - r20], -1)
    q3 = torch.stack([r02 + r20, r21 + r12, (1 + r22) - (r00 + r11), r10 - r01], -1)
    c0 = (trace > 0).unsqueeze(-1)
    c1 = ((r00 > r11) & (r00 > r22)).unsqueeze(-1)
    c2 = (r11 > r22).unsqueeze(-1)
    q = torch.where(c0, q0, torch.where(c1, q1, torch.where(c2, q2, q3)))
    xyz, w = q[..., :3], q[..., 3:]
    n = torch.linalg.norm(xyz, dim=-1, keepdim=True)
    k = torch.where(n == 0, torch.zeros_like(n), 2.0 / torch.where(n == 0, torch.ones_like(n), n))
    return k * torch.atan2(n, w) * xyz
