"""``torch.library`` custom ops that make the modules scriptable (``torch.jit.script``) and
traceable: TorchScript cannot see through ctypes, so the scripted ``forward`` / ``fit`` /
``convert_vertices`` call these ops, whose bodies look the Python module up in a registry and run
the regular host code (pointer marshalling into the C ABI).  Reference behaviour being matched:
tests/conftest.py:38-39 and pt/__init__.py:90 script the fitter.
"""

from __future__ import annotations

import weakref
from typing import List, Optional

import torch

_registry: dict = {}
_next_handle = [1]


def register(module) -> int:
    h = _next_handle[0]
    _next_handle[0] += 1
    _registry[h] = weakref.ref(module)
    return h


def _get(handle: int):
    ref = _registry.get(handle)
    mod = ref() if ref is not None else None
    if mod is None:
        raise RuntimeError(
            f'smplfitter_b200: module handle {handle} is not alive in this process. A scripted smplfitter_b200 module '
            'dispatches to the Python module it was scripted from: keep that module alive, and do not torch.jit.save / '
            'load it into another process (script the module again there).')
    return mod


class RegisteredModule:
    """Mixin of the nn.Modules that dispatch through the ops below: a copy (``copy.deepcopy``, pickling) gets its OWN
    handle, so it never runs on the buffers of the module it was copied from."""

    def __deepcopy__(self, memo):
        import copy

        new = type(self).__new__(type(self))
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        new.__dict__['_handle'] = register(new)
        return new

    def __setstate__(self, state):
        super().__setstate__(state)
        self.__dict__['_handle'] = register(self)


def _empty_like_dev(t: torch.Tensor) -> torch.Tensor:
    return torch.empty(0, device=t.device, dtype=torch.float32)


def _distinct(tensors, inputs):
    """Custom-op contract: returns may alias neither each other nor the inputs."""
    seen = {x.untyped_storage().data_ptr() for x in inputs if x is not None and x.numel() > 0}
    out = []
    for t in tensors:
        if t.numel() > 0:
            p = t.untyped_storage().data_ptr()
            if p in seen:
                t = t.clone()
            seen.add(t.untyped_storage().data_ptr())
        out.append(t)
    return out


@torch.library.custom_op('smplfit_b200::forward', mutates_args=())
def forward_op(handle: int, anchor: torch.Tensor, pose_rotvecs: Optional[torch.Tensor],
               shape_betas: Optional[torch.Tensor], trans: Optional[torch.Tensor],
               kid_factor: Optional[torch.Tensor], rel_rotmats: Optional[torch.Tensor],
               glob_rotmats: Optional[torch.Tensor], return_vertices: bool) -> List[torch.Tensor]:
    res = _get(handle)._forward_impl(pose_rotvecs, shape_betas, trans, kid_factor, rel_rotmats, glob_rotmats,
                                     return_vertices)
    return _distinct([res['joints'], res['orientations'], res['vertices'] if return_vertices else _empty_like_dev(anchor)],
                     [pose_rotvecs, shape_betas, trans, kid_factor, rel_rotmats, glob_rotmats])


@forward_op.register_fake
def _(handle, anchor, pose_rotvecs, shape_betas, trans, kid_factor, rel_rotmats, glob_rotmats, return_vertices):
    m = _get(handle)
    B = 0
    for a in (pose_rotvecs, shape_betas, trans, rel_rotmats, glob_rotmats):
        if a is not None:
            B = a.shape[0]
            break
    new = lambda *s: anchor.new_empty(s, dtype=torch.float32)  # noqa: E731
    return [new(B, m.num_joints, 3), new(B, m.num_joints, 3, 3),
            new(B, m.num_vertices, 3) if return_vertices else new(0)]


def _needs(xs):
    return [x is not None and x.requires_grad for x in xs]


def _forward_setup(ctx, inputs, output):
    handle, _anchor, *tensors, return_vertices = inputs
    ctx.handle, ctx.return_vertices, ctx.needs = handle, return_vertices, _needs(tensors)
    ctx.save_for_backward(*tensors)


def _forward_backward(ctx, grads):
    """Gradient of the forward pass: see ``_adjoint`` (values from CUDA, gradient by a sliced re-evaluation)."""
    from . import _adjoint

    g = _adjoint.forward_backward(_get(ctx.handle), list(ctx.saved_tensors), ctx.needs, ctx.return_vertices, grads)
    return (None, None, *g, None)


forward_op.register_autograd(_forward_backward, setup_context=_forward_setup)


@torch.library.custom_op('smplfit_b200::fit', mutates_args=())
def fit_op(handle: int, target_vertices: torch.Tensor, target_joints: Optional[torch.Tensor],
           vertex_weights: Optional[torch.Tensor], joint_weights: Optional[torch.Tensor], num_iter: int,
           beta_regularizer: float, beta_regularizer2: float, scale_regularizer: float, kid_regularizer: float,
           share_beta: bool, final_adjust_rots: bool, scale_target: bool, scale_fit: bool,
           initial_pose_rotvecs: Optional[torch.Tensor], initial_shape_betas: Optional[torch.Tensor],
           initial_kid_factor: Optional[torch.Tensor], want_pose_rotvecs: bool,
           want_rel_orient: bool) -> List[torch.Tensor]:
    keys = (['pose_rotvecs'] if want_pose_rotvecs else []) + (['relative_orientations'] if want_rel_orient else [])
    res = _get(handle)._fit_impl(
        target_vertices, target_joints, vertex_weights, joint_weights, num_iter, beta_regularizer, beta_regularizer2,
        scale_regularizer, None if kid_regularizer != kid_regularizer else kid_regularizer, share_beta,
        final_adjust_rots, scale_target, scale_fit, initial_pose_rotvecs, initial_shape_betas, initial_kid_factor, keys)
    e = lambda: _empty_like_dev(target_vertices)  # noqa: E731
    return _distinct(
        [res['shape_betas'], res['trans'], res['orientations'], res['relative_orientations'],
         res['pose_rotvecs'] if 'pose_rotvecs' in res else e(), res['kid_factor'] if 'kid_factor' in res else e(),
         res['scale_corr'] if 'scale_corr' in res else e()],
        [target_vertices, target_joints, vertex_weights, joint_weights, initial_pose_rotvecs, initial_shape_betas,
         initial_kid_factor])


@fit_op.register_fake
def _(handle, target_vertices, target_joints, vertex_weights, joint_weights, num_iter, beta_regularizer,
      beta_regularizer2, scale_regularizer, kid_regularizer, share_beta, final_adjust_rots, scale_target, scale_fit,
      initial_pose_rotvecs, initial_shape_betas, initial_kid_factor, want_pose_rotvecs, want_rel_orient):
    f = _get(handle)
    B, J = target_vertices.shape[0], f.body_model.num_joints
    new = lambda *s: target_vertices.new_empty(s, dtype=torch.float32)  # noqa: E731
    return [new(B, f.n_betas), new(B, 3), new(B, J, 3, 3), new(B, J, 3, 3),
            new(B, 3 * J) if want_pose_rotvecs else new(0), new(B) if f.enable_kid else new(0),
            new(B) if (scale_target or scale_fit) else new(0)]


def _fit_setup(ctx, inputs, output):
    (handle, tv, tj, vw, jw, num_iter, reg, reg2, scale_reg, kid_reg, share_beta, final_adjust_rots, scale_target,
     scale_fit, init_pose, init_betas, init_kid, want_rv, want_rel) = inputs
    tensors = [tv, tj, vw, jw, init_pose, init_betas, init_kid]
    ctx.handle, ctx.needs = handle, _needs(tensors)
    ctx.opts = dict(num_iter=num_iter, beta_regularizer=reg, beta_regularizer2=reg2, scale_regularizer=scale_reg,
                    kid_regularizer=None if kid_reg != kid_reg else kid_reg, share_beta=share_beta,
                    final_adjust_rots=final_adjust_rots, scale_target=scale_target, scale_fit=scale_fit,
                    want_pose_rotvecs=want_rv, want_rel_orient=want_rel)
    ctx.save_for_backward(*tensors)


def _fit_backward(ctx, grads):
    """Gradient of the fit with respect to targets, weights and initial guesses: see ``_adjoint``."""
    from . import _adjoint

    g = _adjoint.fit_backward(_get(ctx.handle), list(ctx.saved_tensors), ctx.needs, ctx.opts, grads)
    return (None, g[0], g[1], g[2], g[3], None, None, None, None, None, None, None, None, None, g[4], g[5], g[6], None,
            None)


fit_op.register_autograd(_fit_backward, setup_context=_fit_setup)


@torch.library.custom_op('smplfit_b200::convert_vertices', mutates_args=())
def convert_vertices_op(handle: int, inp_vertices: torch.Tensor) -> torch.Tensor:
    return _distinct([_get(handle)._convert_vertices_impl(inp_vertices)], [inp_vertices])[0]


@convert_vertices_op.register_fake
def _(handle, inp_vertices):
    c = _get(handle)
    return inp_vertices.new_empty((inp_vertices.shape[0], c.body_model_out.num_vertices, 3), dtype=torch.float32)


def _convert_setup(ctx, inputs, output):
    ctx.handle = inputs[0]


def _convert_backward(ctx, grad):
    """The topology transfer is linear (out = M x per coordinate): the pull-back is M^T, applied as a scatter-add over
    the non-zeros (pt/bodyconverter.py:129-149 is differentiable the same way through its sparse matmul)."""
    return None, _get(ctx.handle)._convert_vertices_transpose(grad)


convert_vertices_op.register_autograd(_convert_backward, setup_context=_convert_setup)
