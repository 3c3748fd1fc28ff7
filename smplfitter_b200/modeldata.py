"""Model constants for the SMPL-family fitter and a synthetic SMPL-shaped model generator.

``ModelData`` mirrors the 15 fields of the reference container
(/root/reference/src/smplfitter/common.py:156-209) so a reference-style loader can be
plugged into ``BodyModel`` unchanged.  The licensed SMPL / SMPL-X files are not
redistributable and are absent from the build and GPU boxes, so all parity and benchmark
work runs on a *synthetic* model with the real dimensions and kinematic trees
(SURVEY.md section 8c).  The derivations applied after sampling (joint regressor ->
``J_template`` / ``J_shapedirs``, pose-corrected ``v_template``) follow
common.py:336-350.
"""

from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
SMPLX_PARENTS = SMPL_PARENTS[:22] + [
    15, 15, 15,
    20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
    21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53,
]

# Approximate rest-pose joint layout of a standing human (metres, y up, x to the body's left).
_SMPL_REST = np.array(
    [
        [0.00, -0.24, 0.03], [0.07, -0.33, 0.02], [-0.07, -0.33, 0.02], [0.00, -0.11, 0.00],
        [0.10, -0.71, 0.02], [-0.10, -0.71, 0.02], [0.00, 0.02, 0.02], [0.09, -1.11, -0.02],
        [-0.09, -1.11, -0.02], [0.00, 0.08, 0.03], [0.11, -1.17, 0.10], [-0.11, -1.17, 0.10],
        [0.00, 0.29, -0.01], [0.08, 0.20, 0.00], [-0.08, 0.20, 0.00], [0.00, 0.38, 0.04],
        [0.17, 0.23, -0.01], [-0.17, 0.23, -0.01], [0.43, 0.22, -0.03], [-0.43, 0.22, -0.03],
        [0.68, 0.23, -0.03], [-0.68, 0.23, -0.03], [0.76, 0.22, -0.04], [-0.76, 0.22, -0.04],
    ],
    dtype=np.float64,
)


@dataclass
class ModelData:
    """Arrays and metadata of one body model (float64 numpy, cast to float32 by BodyModel)."""

    v_template: np.ndarray  # (V, 3), pose-corrected (common.py:346-350)
    shapedirs: np.ndarray  # (V, 3, S)
    posedirs: np.ndarray  # (V, 3, 9*(J-1))
    J_regressor_post_lbs: np.ndarray  # (J, V)
    J_template: np.ndarray  # (J, 3)
    J_shapedirs: np.ndarray  # (J, 3, S)
    kid_shapedir: np.ndarray  # (V, 3)
    kid_J_shapedir: np.ndarray  # (J, 3)
    weights: np.ndarray  # (V, J)
    kintree_parents: list
    faces: np.ndarray
    num_joints: int
    num_vertices: int
    vertex_subset: np.ndarray
    joint_names: list = field(default_factory=list)


def _rest_joints(parents: list, rng: np.random.RandomState) -> np.ndarray:
    """Rest joint positions: anatomical table for the SMPL / SMPL-X trees, random walk otherwise."""
    J = len(parents)
    if parents == SMPL_PARENTS:
        return _SMPL_REST.copy()
    if parents == SMPLX_PARENTS:
        j = np.zeros((J, 3))
        j[:22] = _SMPL_REST[:22]
        j[22] = j[15] + [0.0, -0.02, 0.06]  # jaw
        j[23] = j[15] + [0.03, 0.05, 0.08]  # eyes
        j[24] = j[15] + [-0.03, 0.05, 0.08]
        for side, wrist, start in ((1.0, 20, 25), (-1.0, 21, 40)):
            for f in range(5):
                spread = (f - 2) * 0.018
                base = j[wrist] + [side * 0.09, 0.0, spread]
                for k in range(3):
                    j[start + 3 * f + k] = base + [side * 0.03 * k, -0.004 * k, 0.3 * spread * k]
        return j
    j = np.zeros((J, 3))
    for i in range(1, J):
        d = rng.randn(3)
        j[i] = j[parents[i]] + 0.15 * d / np.linalg.norm(d)
    return j


def make_synthetic_model_data(
    parents: Optional[list] = None,
    num_vertices: int = 6890,
    num_betas: int = 10,
    seed: int = 0,
    max_influences: int = 4,
    with_kid: bool = False,
) -> ModelData:
    """Sample a synthetic body model with SMPL-like structure.

    Every body part is an anisotropic (elliptic) tube around the segment from its joint
    towards the mean of its child joints, so bone twist is observable; skin weights are a
    distance soft-max truncated to ``max_influences`` joints per vertex (real SMPL has at
    most 4 non-zeros per row); vertex indices are coherent per part with local shuffling,
    like a real mesh numbering.  ``J_regressor`` rows sum to one.
    """
    if parents is None:
        parents = SMPL_PARENTS
    rng = np.random.RandomState(seed)
    J = len(parents)
    V = int(num_vertices)
    S = int(num_betas)
    P = 9 * (J - 1)
    joints = _rest_joints(list(parents), rng)

    children = [[] for _ in range(J)]
    for i in range(1, J):
        children[parents[i]].append(i)
    seg_end = np.zeros((J, 3))
    for i in range(J):
        if children[i]:
            seg_end[i] = joints[children[i]].mean(axis=0)
        elif i > 0:
            d = joints[i] - joints[parents[i]]
            dn = np.linalg.norm(d)
            seg_end[i] = joints[i] + d / dn * max(0.8 * dn, 0.12 if J <= 24 or i < 22 else 0.025)
        else:
            seg_end[i] = joints[i] + [0.0, 0.1, 0.0]
    seg_len = np.linalg.norm(seg_end - joints, axis=1)

    # Vertex ownership: part sizes grow with segment length, each part gets a floor share.
    share = 0.35 / J + 0.65 * seg_len / seg_len.sum()
    counts = np.maximum(np.floor(share * V).astype(np.int64), min(8, V // J))
    while counts.sum() > V:
        counts[np.argmax(counts)] -= 1
    counts[np.argmax(share)] += V - counts.sum()
    owner = np.repeat(np.arange(J), counts)
    # local shuffles keep the numbering coherent but not sorted by part
    for _ in range(V // 8):
        a = rng.randint(0, V)
        b = min(V - 1, max(0, a + rng.randint(-48, 49)))
        owner[a], owner[b] = owner[b], owner[a]

    axis = seg_end - joints
    axis /= np.maximum(np.linalg.norm(axis, axis=1, keepdims=True), 1e-9)
    helper = np.where(np.abs(axis[:, [1]]) < 0.9, [[0.0, 1.0, 0.0]], [[1.0, 0.0, 0.0]])
    e1 = np.cross(axis, helper)
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(axis, e1)
    radius = np.clip(0.45 * seg_len, 0.035 if J <= 24 else 0.012, 0.11)
    u = rng.uniform(-0.1, 1.1, size=V)
    phi = rng.uniform(0, 2 * np.pi, size=V)
    rr = np.sqrt(rng.uniform(0.3, 1.0, size=V))
    o = owner
    v_orig = (
        joints[o]
        + (u * seg_len[o])[:, None] * axis[o]
        + (rr * radius[o] * 1.0 * np.cos(phi))[:, None] * e1[o]
        + (rr * radius[o] * 0.6 * np.sin(phi))[:, None] * e2[o]
    )

    # Skin weights: soft-max of squared distance to each part's segment, owner boosted.
    ab = seg_end - joints
    ab2 = np.maximum((ab * ab).sum(axis=1), 1e-12)
    d2 = np.empty((V, J))
    for j in range(J):
        tpar = np.clip(((v_orig - joints[j]) @ ab[j]) / ab2[j], 0.0, 1.0)
        closest = joints[j] + tpar[:, None] * ab[j]
        d2[:, j] = ((v_orig - closest) ** 2).sum(axis=1)
    logits = -d2 / 0.004
    logits[np.arange(V), owner] += 1.5
    logits -= logits.max(axis=1, keepdims=True)
    w = np.exp(logits)
    if max_influences < J:
        keep = np.argsort(-w, axis=1, kind='stable')[:, :max_influences]
        mask = np.zeros_like(w, dtype=bool)
        mask[np.arange(V)[:, None], keep] = True
        w = np.where(mask, w, 0.0)
    w /= w.sum(axis=1, keepdims=True)

    J_regressor = w.T / w.sum(axis=0, keepdims=True).T
    shapedirs = 0.01 * rng.randn(V, 3, S)
    posedirs = 0.002 * rng.randn(V, 3, P)

    if with_kid:
        v_smil = 0.6 * (v_orig - v_orig.mean(axis=0)) + 0.004 * rng.randn(V, 3)
        kid_shapedir = v_smil - v_smil.mean(axis=0) - v_orig
        kid_J_shapedir = J_regressor @ kid_shapedir
    else:
        kid_shapedir = np.zeros((V, 3))
        kid_J_shapedir = np.zeros((J, 3))

    # Same derivations as common.py:336-350.
    J_shapedirs = np.einsum('jv,vcs->jcs', J_regressor, shapedirs)
    J_template = J_regressor @ v_orig
    eye_feat = np.reshape(np.tile(np.eye(3), [J - 1, 1]), [-1])
    v_template = v_orig - np.einsum('vcx,x->vc', posedirs, eye_feat)

    return ModelData(
        v_template=v_template,
        shapedirs=shapedirs,
        posedirs=posedirs,
        J_regressor_post_lbs=J_regressor,
        J_template=J_template,
        J_shapedirs=J_shapedirs,
        kid_shapedir=kid_shapedir,
        kid_J_shapedir=kid_J_shapedir,
        weights=w,
        kintree_parents=list(parents),
        faces=np.zeros((0, 3), dtype=np.int32),
        num_joints=J,
        num_vertices=V,
        vertex_subset=np.arange(V, dtype=np.int64),
        joint_names=[f'joint{i}' for i in range(J)],
    )


def apply_vertex_subset(
    data: ModelData, vertex_subset, joint_regressor_post_lbs=None, faces=None
) -> ModelData:
    """Restrict a model to a vertex subset (common.py:368-394 semantics)."""
    vs = np.asarray(vertex_subset, dtype=np.int64)
    if joint_regressor_post_lbs is None:
        joint_regressor_post_lbs = data.J_regressor_post_lbs[:, vs]
    return ModelData(
        v_template=data.v_template[vs],
        shapedirs=data.shapedirs[vs],
        posedirs=data.posedirs[vs],
        J_regressor_post_lbs=np.asarray(joint_regressor_post_lbs),
        J_template=data.J_template,
        J_shapedirs=data.J_shapedirs,
        kid_shapedir=data.kid_shapedir[vs],
        kid_J_shapedir=data.kid_J_shapedir,
        weights=data.weights[vs],
        kintree_parents=data.kintree_parents,
        faces=data.faces if faces is None else faces,
        num_joints=data.num_joints,
        num_vertices=len(vs),
        vertex_subset=vs,
        joint_names=data.joint_names,
    )


# name -> (parents, V, S_full, seed, with_kid): the synthetic stand-ins for the licensed models.
SYNTHETIC_SPECS = {
    'smpl': (SMPL_PARENTS, 6890, 10, 0, True),
    'smplx': (SMPLX_PARENTS, 10475, 16, 1, True),
    'smpl_tiny': (SMPL_PARENTS, 431, 10, 2, True),
    'smplx_tiny': (SMPLX_PARENTS, 977, 16, 3, True),
}

_provider: Optional[Callable[..., ModelData]] = None
_cache: dict = {}
_synthetic_enabled: Optional[bool] = None  # None: follow $SMPLFITTER_B200_SYNTHETIC (default off)


def use_synthetic_models(enable: bool = True) -> None:
    """Opt in to (or out of) the synthetic stand-ins for model names without a provider.

    Off by default: like the reference (common.py:219-260), ``BodyModel('smpl')`` raises
    ``FileNotFoundError`` when the licensed files cannot be loaded.  Tests, ``bench.py`` and
    ``smoke()`` call this explicitly; ``SMPLFITTER_B200_SYNTHETIC=1`` does the same from the
    environment."""
    global _synthetic_enabled
    _synthetic_enabled = bool(enable)


def synthetic_models_enabled() -> bool:
    if _synthetic_enabled is not None:
        return _synthetic_enabled
    return os.environ.get('SMPLFITTER_B200_SYNTHETIC', '0') == '1'


def set_model_provider(fn: Optional[Callable[..., ModelData]]) -> None:
    """Install a loader ``fn(model_name, gender, model_root, num_betas, vertex_subset_size,
    vertex_subset, faces, joint_regressor_post_lbs) -> ModelData`` (e.g. the reference's
    ``smplfitter.common.initialize`` when the licensed files are available)."""
    global _provider
    _provider = fn


def synthetic_model(name: str) -> ModelData:
    if name not in _cache:
        parents, V, S, seed, kid = SYNTHETIC_SPECS[name]
        _cache[name] = make_synthetic_model_data(parents, V, S, seed, with_kid=kid)
    return _cache[name]


def initialize(
    model_name,
    gender='neutral',
    model_root=None,
    num_betas=None,
    vertex_subset_size=None,
    vertex_subset=None,
    faces=None,
    joint_regressor_post_lbs=None,
) -> ModelData:
    """Same call signature as the reference seam (common.py:219-228).

    Resolution order: an installed provider; else, only after an explicit opt-in
    (``use_synthetic_models()`` or ``SMPLFITTER_B200_SYNTHETIC=1``), the synthetic stand-in of
    that name; else ``FileNotFoundError`` as in the reference.  Reading the licensed files is
    delegated to a provider on purpose (file parsing is out of scope).
    """
    if _provider is not None:
        return _provider(
            model_name, gender, model_root, num_betas, vertex_subset_size, vertex_subset,
            faces, joint_regressor_post_lbs,
        )
    if not synthetic_models_enabled() or model_name not in SYNTHETIC_SPECS:
        raise FileNotFoundError(
            f"No model provider installed for '{model_name}'. Call "
            'smplfitter_b200.modeldata.set_model_provider(smplfitter.common.initialize) to load '
            'the licensed files; synthetic stand-ins (' + ', '.join(SYNTHETIC_SPECS) + ') are only '
            'served after smplfitter_b200.modeldata.use_synthetic_models() or with '
            'SMPLFITTER_B200_SYNTHETIC=1'
        )
    data = synthetic_model(model_name)
    if vertex_subset_size is not None and vertex_subset is None:
        vertex_subset = np.sort(
            np.random.RandomState(42).choice(data.num_vertices, vertex_subset_size, replace=False)
        )
    if vertex_subset is not None:
        data = apply_vertex_subset(data, vertex_subset, joint_regressor_post_lbs, faces)
    elif joint_regressor_post_lbs is not None:
        data = apply_vertex_subset(
            data, np.arange(data.num_vertices), joint_regressor_post_lbs, faces
        )
    if num_betas is not None:
        data = ModelData(**{**data.__dict__})
        data.shapedirs = data.shapedirs[:, :, :num_betas]
        data.J_shapedirs = data.J_shapedirs[:, :, :num_betas]
    return data
