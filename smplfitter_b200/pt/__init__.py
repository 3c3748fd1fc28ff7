"""Drop-in surface of ``smplfitter.pt`` (/root/reference/src/smplfitter/pt/__init__.py:25-32)."""

from __future__ import annotations

import functools

from .bodyconverter import BodyConverter
from .bodyfitter import BodyFitter
from .bodyflipper import BodyFlipper
from .bodymodel import BodyModel

__all__ = ['BodyModel', 'BodyFitter', 'BodyConverter', 'BodyFlipper', 'get_cached_body_model', 'get_cached_fit_fn']


@functools.lru_cache()
def get_cached_body_model(model_name='smpl', gender='neutral', model_root=None):
    """Cached ``BodyModel`` (pt/__init__.py:36-55)."""
    return BodyModel(model_root=model_root, gender=gender, model_name=model_name)


@functools.lru_cache()
def get_cached_fit_fn(
    body_model_name='smpl', gender='neutral', num_betas=10, enable_kid=False,
    requested_keys=('pose_rotvecs', 'shape_betas', 'trans'), beta_regularizer=1.0, beta_regularizer2=0.0,
    num_iter=3, vertex_subset=None, joint_regressor_post_lbs=None, share_beta=False,
    final_adjust_rots=True, scale_target=False, scale_fit=False, scale_regularizer=0.0,
    kid_regularizer=None, device='cuda',
):
    """Closure over a fitter with baked options, flattening leading dims (pt/__init__.py:58-132)."""
    body_model = BodyModel(
        gender=gender, model_name=body_model_name, num_betas=num_betas, vertex_subset=vertex_subset,
        joint_regressor_post_lbs=joint_regressor_post_lbs, device=device,
    )
    fitter = BodyFitter(body_model, enable_kid=enable_kid).to(device)

    def wrapped(verts, joints=None, vertex_weights=None, joint_weights=None):
        lead = verts.shape[:-2]
        flat = lambda x, *tail: x.reshape(-1, *tail) if x is not None else None  # noqa: E731
        res = fitter.fit(
            flat(verts, body_model.num_vertices, 3), flat(joints, body_model.num_joints, 3),
            flat(vertex_weights, body_model.num_vertices), flat(joint_weights, body_model.num_joints),
            num_iter=num_iter, beta_regularizer=beta_regularizer, beta_regularizer2=beta_regularizer2,
            scale_regularizer=scale_regularizer, kid_regularizer=kid_regularizer, share_beta=share_beta,
            final_adjust_rots=final_adjust_rots, scale_target=scale_target, scale_fit=scale_fit,
            requested_keys=list(requested_keys),
        )
        return {k: v.reshape(*lead, *v.shape[1:]) for k, v in res.items()}

    return wrapped
