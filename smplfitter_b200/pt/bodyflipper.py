"""``BodyFlipper`` -- drop-in for ``smplfitter.pt.BodyFlipper``
(/root/reference/src/smplfitter/pt/bodyflipper.py:18-130): mirrors a body along the x axis by flipping and
re-ordering the vertices (sparse mirror transfer) and fitting the parameters to the flipped mesh, starting from
the naively flipped pose.  Forward LBS, the sparse transfer and the fit all run in the CUDA library.
"""

from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _native
from .bodyfitter import BodyFitter


def get_mirror_mapping(points: np.ndarray) -> np.ndarray:
    """Index of the mirror partner of every point (pt/bodyflipper.py:133-137: optimal assignment between the
    points and their x-flipped copies)."""
    import scipy.optimize
    import scipy.spatial.distance

    pts = np.asarray(points, np.float64)
    dist = scipy.spatial.distance.cdist(pts, pts * [-1, 1, 1])
    v_inds, mirror_inds = scipy.optimize.linear_sum_assignment(dist)
    return mirror_inds[np.argsort(v_inds)].astype(np.int64)


def nearest_mirror_csr(vertices: np.ndarray):
    """Stand-in mirror transfer for models without the licensed SMPL-X flip correspondences: every vertex takes the
    position of the vertex nearest to its mirror image (a one-non-zero-per-row CSR matrix)."""
    import scipy.sparse
    import scipy.spatial

    v = np.asarray(vertices, np.float64)
    _, nn_idx = scipy.spatial.cKDTree(v).query(v * [-1, 1, 1])
    n = v.shape[0]
    return scipy.sparse.csr_matrix((np.ones(n, np.float32), (np.arange(n), nn_idx)), shape=(n, n))


class BodyFlipper(nn.Module):
    """Horizontally (x axis) flips SMPL-like body model parameters.

    ``mirror_csr`` may be given explicitly (scipy CSR, (V, V)).  Otherwise, as in the reference (:140-156), the
    SMPL-X flip correspondences (and, for SMPL, the two transfer matrices) are read from
    ``$DATA_ROOT/body_models``; without those licensed files the constructor raises ``FileNotFoundError`` like the
    reference, unless synthetic models were opted into (``modeldata.use_synthetic_models()``), in which case the
    nearest-mirror-vertex stand-in is used.
    """

    def __init__(self, body_model, mirror_csr=None):
        super().__init__()
        self.body_model = body_model
        self.fitter = BodyFitter(self.body_model, enable_kid=True)
        res = {k: v.detach().cpu().numpy() for k, v in self._rest_pose().items()}
        if mirror_csr is None:
            mirror_csr = self._load_default_csr(body_model.num_vertices)
        if mirror_csr is None:
            from .. import modeldata

            if not modeldata.synthetic_models_enabled():
                raise FileNotFoundError(
                    'smplx_flip_correspondences.npz not found under $DATA_ROOT/body_models/smplx; pass mirror_csr= '
                    'explicitly (the nearest-mirror stand-in is only used with synthetic models)')
            mirror_csr = nearest_mirror_csr(res['vertices'])
        m = mirror_csr.tocsr().astype(np.float32)
        V = body_model.num_vertices
        if m.shape != (V, V):
            raise ValueError(f'mirror_csr must be {(V, V)}')
        self.register_buffer('_csr_indptr', torch.tensor(m.indptr, dtype=torch.int32), persistent=False)
        self.register_buffer('_csr_indices', torch.tensor(m.indices, dtype=torch.int32), persistent=False)
        self.register_buffer('_csr_data', torch.tensor(m.data, dtype=torch.float32), persistent=False)
        self.mirror_inds_joints = nn.Buffer(torch.tensor(get_mirror_mapping(res['joints'])))
        self.to(body_model.v_template.device)

    def _rest_pose(self):
        """Zero-pose, zero-shape vertices and joints (``body_model.single()`` of the reference) from the model
        constants: no device needed at construction time."""
        bm = self.body_model
        return dict(vertices=bm._t_template_mesh, joints=bm.J_template)

    @staticmethod
    def _load_default_csr(num_verts):
        data_root = os.getenv('DATA_ROOT', '.')
        flip_path = f'{data_root}/body_models/smplx/smplx_flip_correspondences.npz'
        if not os.path.exists(flip_path):
            return None
        import pickle

        import scipy.sparse

        mfile = np.load(flip_path)
        faces, bc = mfile['closest_faces'], mfile['bc']
        smplx2mirror = scipy.sparse.coo_matrix(
            (bc.flatten(), (np.repeat(np.arange(faces.shape[0]), 3), faces.flatten())),
            shape=(faces.shape[0], bc.shape[0])).tocsr().astype(np.float32)

        def transfer(name):
            with open(f'{data_root}/body_models/{name}', 'rb') as f:
                mtx = pickle.load(f, encoding='latin1')['mtx'].tocsr().astype(np.float32)
            return mtx[:, : mtx.shape[1] // 2]

        if num_verts == 10475:
            return smplx2mirror
        if num_verts == 6890:
            return transfer('smplx2smpl_deftrafo_setup.pkl') @ smplx2mirror @ transfer('smpl2smplx_deftrafo_setup.pkl')
        raise ValueError(f'Unsupported number of vertices: {num_verts}')

    def flip(self, pose_rotvecs: torch.Tensor, shape_betas: torch.Tensor, trans: torch.Tensor,
             kid_factor: Optional[torch.Tensor] = None, num_iter: int = 1) -> dict:
        """Parameters of the horizontally flipped body (pt/bodyflipper.py:36-89)."""
        inp = self.body_model(pose_rotvecs, shape_betas, trans, kid_factor=kid_factor)
        flipped_vertices = self.flip_vertices(inp['vertices'])
        fit = self.fitter.fit(
            target_vertices=flipped_vertices, num_iter=num_iter, beta_regularizer=1e-2, beta_regularizer2=1e-2,
            final_adjust_rots=True, kid_regularizer=1e9 if kid_factor is None else 0.0,
            initial_pose_rotvecs=self.naive_flip_rotvecs(pose_rotvecs), initial_shape_betas=shape_betas,
            requested_keys=['pose_rotvecs', 'shape_betas'],
        )
        return dict(pose_rotvecs=fit['pose_rotvecs'], shape_betas=fit['shape_betas'], trans=fit['trans'],
                    kid_factor=fit.get('kid_factor'))

    def flip_vertices(self, inp_vertices: torch.Tensor) -> torch.Tensor:
        """Mirror transfer then x flip (pt/bodyflipper.py:91-110); the CSR product runs in ``smplfit_convert_vertices``."""
        _native.require_cuda(self._csr_data, 'the flipper')
        dev = self._csr_data.device
        x = inp_vertices.to(device=dev, dtype=torch.float32).contiguous()
        B, V = x.shape[0], self.body_model.num_vertices
        out = torch.empty((B, V, 3), device=dev, dtype=torch.float32)
        if B > 0:
            with torch.cuda.device(dev):
                _native.check(_native.lib().smplfit_convert_vertices(
                    self._csr_indptr.data_ptr(), self._csr_indices.data_ptr(), self._csr_data.data_ptr(),
                    V, V, B, x.data_ptr(), out.data_ptr(), _native.stream_ptr(dev),
                ))
        return out * torch.tensor([-1.0, 1.0, 1.0], dtype=out.dtype, device=dev)

    def naive_flip_rotvecs(self, pose_rotvecs: torch.Tensor) -> torch.Tensor:
        """Swap left / right parts and mirror each rotation vector (pt/bodyflipper.py:112-130)."""
        mult = torch.tensor([1, -1, -1], dtype=pose_rotvecs.dtype, device=pose_rotvecs.device)
        J = self.body_model.num_joints
        reshaped = pose_rotvecs.reshape(-1, J, 3)
        return (reshaped[:, self.mirror_inds_joints.to(pose_rotvecs.device)] * mult).reshape(-1, J * 3)
