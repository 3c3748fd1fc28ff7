"""``BodyConverter`` -- drop-in for ``smplfitter.pt.BodyConverter``
(/root/reference/src/smplfitter/pt/bodyconverter.py:14-158): forward LBS of the input model,
sparse topology transfer, fit of the output model -- all three steps in the CUDA library.
"""

from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _native
from . import _ops
from .bodyfitter import BodyFitter


class BodyConverter(_ops.RegisteredModule, nn.Module):
    """Converts between SMPL-family parametrisations.

    ``vertex_converter_csr`` may be passed explicitly as a ``scipy.sparse`` CSR matrix of
    shape (V_out, V_in); otherwise, as in the reference (:32-46), it is looked up under
    ``$DATA_ROOT/body_models`` for the SMPL<->SMPL-X pair through an installed loader and
    left ``None`` (identity topology) when the models share their mesh.
    """

    def __init__(self, body_model_in, body_model_out, vertex_converter_csr=None):
        super().__init__()
        self.body_model_in = body_model_in
        self.body_model_out = body_model_out
        self.fitter = BodyFitter(self.body_model_out, enable_kid=True)
        if vertex_converter_csr is None and body_model_in.num_vertices != body_model_out.num_vertices:
            vertex_converter_csr = self._load_default_csr()
        self.has_converter = vertex_converter_csr is not None
        if self.has_converter:
            m = vertex_converter_csr.tocsr().astype(np.float32)
            if m.shape != (body_model_out.num_vertices, body_model_in.num_vertices):
                raise ValueError(f'vertex converter must be {(body_model_out.num_vertices, body_model_in.num_vertices)}')
            self.register_buffer('_csr_indptr', torch.tensor(m.indptr, dtype=torch.int32), persistent=False)
            self.register_buffer('_csr_indices', torch.tensor(m.indices, dtype=torch.int32), persistent=False)
            self.register_buffer('_csr_data', torch.tensor(m.data, dtype=torch.float32), persistent=False)
        self._handle = _ops.register(self)

    @torch.jit.unused
    def _load_default_csr(self):
        data_root = os.getenv('DATA_ROOT', '.')
        vin, vout = self.body_model_in.num_vertices, self.body_model_out.num_vertices
        if vin == 6890 and vout == 10475:
            path = f'{data_root}/body_models/smpl2smplx_deftrafo_setup.pkl'
        elif vin == 10475 and vout == 6890:
            path = f'{data_root}/body_models/smplx2smpl_deftrafo_setup.pkl'
        else:
            return None
        if not os.path.exists(path):
            raise FileNotFoundError(
                f'{path} not found: pass vertex_converter_csr=... explicitly (the licensed transfer '
                'matrices are not redistributable).'
            )
        import pickle

        with open(path, 'rb') as f:
            m = pickle.load(f, encoding='latin1')['mtx'].tocsr().astype(np.float32)
        return m[:, : m.shape[1] // 2]  # common.py:425-429

    @torch.jit.unused
    def convert(
        self,
        pose_rotvecs: torch.Tensor,
        shape_betas: torch.Tensor,
        trans: torch.Tensor,
        kid_factor: Optional[torch.Tensor] = None,
        known_output_pose_rotvecs: Optional[torch.Tensor] = None,
        known_output_shape_betas: Optional[torch.Tensor] = None,
        known_output_kid_factor: Optional[torch.Tensor] = None,
        num_iter: int = 1,
    ) -> dict[str, torch.Tensor]:
        """pt/bodyconverter.py:48-127."""
        inp_vertices = self.body_model_in(pose_rotvecs, shape_betas, trans)['vertices']
        verts = self.convert_vertices(inp_vertices)
        if known_output_shape_betas is not None:
            fit = self.fitter.fit_with_known_shape(
                shape_betas=known_output_shape_betas, kid_factor=known_output_kid_factor,
                target_vertices=verts, num_iter=num_iter, final_adjust_rots=False,
                requested_keys=['pose_rotvecs'],
            )
            return dict(pose_rotvecs=fit['pose_rotvecs'], trans=fit['trans'])
        if known_output_pose_rotvecs is not None:
            fit = self.fitter.fit_with_known_pose(
                pose_rotvecs=known_output_pose_rotvecs, target_vertices=verts, beta_regularizer=0.0,
                kid_regularizer=1e9 if kid_factor is None else 0.0,
            )
            fit_out = dict(shape_betas=fit['shape_betas'], trans=fit['trans'])
        else:
            fit = self.fitter.fit(
                target_vertices=verts, num_iter=num_iter, beta_regularizer=0.0, final_adjust_rots=False,
                kid_regularizer=1e9 if kid_factor is None else 0.0,
                requested_keys=['pose_rotvecs', 'shape_betas'],
            )
            fit_out = dict(pose_rotvecs=fit['pose_rotvecs'], shape_betas=fit['shape_betas'], trans=fit['trans'])
        if kid_factor is not None:
            fit_out['kid_factor'] = fit['kid_factor']
        return fit_out

    @torch.jit.export
    def convert_vertices(self, inp_vertices: torch.Tensor) -> torch.Tensor:
        """Barycentric topology transfer (pt/bodyconverter.py:129-149)."""
        if not self.has_converter:
            return inp_vertices
        return torch.ops.smplfit_b200.convert_vertices(self._handle, inp_vertices)

    @torch.jit.unused
    def _convert_vertices_transpose(self, grad_out: torch.Tensor) -> torch.Tensor:
        """(B, V_out, 3) cotangent -> (B, V_in, 3): M^T applied non-zero by non-zero (backward of ``convert_vertices``)."""
        dev = grad_out.device
        indptr = self._csr_indptr.to(dev).long()
        cols, data = self._csr_indices.to(dev).long(), self._csr_data.to(dev)
        rows = torch.repeat_interleave(torch.arange(indptr.numel() - 1, device=dev), indptr[1:] - indptr[:-1])
        B = grad_out.shape[0]
        out = torch.zeros((B, self.body_model_in.num_vertices, 3), device=dev, dtype=grad_out.dtype)
        step = max(1, int(2 ** 28 // max(1, 12 * rows.numel())))  # <= 256 MB of gathered non-zeros at a time
        for a in range(0, B, step):
            g = grad_out[a:a + step]
            out[a:a + step].index_add_(1, cols, g[:, rows] * data.to(g.dtype)[None, :, None])
        return out

    @torch.jit.unused
    def _convert_vertices_impl(self, inp_vertices: torch.Tensor) -> torch.Tensor:
        _native.require_cuda(self._csr_data, 'the converter')
        dev = self._csr_data.device
        x = inp_vertices.to(device=dev, dtype=torch.float32).contiguous()
        B = x.shape[0]
        vin, vout = self.body_model_in.num_vertices, self.body_model_out.num_vertices
        out = torch.empty((B, vout, 3), device=dev, dtype=torch.float32)
        if B == 0:
            return out
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplfit_convert_vertices(
                self._csr_indptr.data_ptr(), self._csr_indices.data_ptr(), self._csr_data.data_ptr(),
                vout, vin, B, x.data_ptr(), out.data_ptr(), _native.stream_ptr(dev),
            ))
        return out
