"""Shared description of the golden fixtures (written by oracle/make_golden.py)."""

import os

import numpy as np

from oracle.make_golden import (CONVERT_CASES, FIT_CASES, FORWARD_CASES, KNOWN_POSE_CASES,  # noqa: F401
                                KNOWN_SHAPE_CASES, MASK_CASES, aux_call_kwargs)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz')))


def fit_call_kwargs(name, g, conv=lambda x: x):
    """Rebuild the exact fit() kwargs of a golden fit case from its fixture."""
    _, _, _, _, _, _, fkw, flags = FIT_CASES[name]
    kw = dict(fkw)
    kw['target_vertices'] = conv(g['target_vertices'])
    if flags.get('joints'):
        kw['target_joints'] = conv(g['target_joints'])
    if 'in_vw' in g:
        kw['vertex_weights'] = conv(g['in_vw'])
    if 'in_jw' in g:
        kw['joint_weights'] = conv(g['in_jw'])
    if 'in_init_pose' in g:
        kw['initial_pose_rotvecs'] = conv(g['in_init_pose'])
        kw['initial_shape_betas'] = conv(g['in_init_betas'])
    kw['requested_keys'] = ['pose_rotvecs', 'shape_betas']
    return kw


def orient_tolerance(g, floor=1e-4, k=6.0):
    """Per-joint tolerance on rotation-matrix entries: the reference's own measured
    reproducibility on this input (vertex renumbering moves its fp32 sums) with a floor."""
    n = g['ref_noise_orient']
    return np.maximum(floor, np.maximum(k * n, n.max()))


def relative_tolerance(g, parents, floor=1e-4, k=6.0):
    """Relative rotation R_parent^T R_j: carries the noise of both joints."""
    t = orient_tolerance(g, floor, k)
    par = np.asarray(parents).copy()
    par[0] = 0
    rel = t + t[par]
    rel[0] = t[0]
    return rel


def rotvec_tolerance(g, parents, floor=1e-4, k=6.0):
    """Rotation-vector entries of the relative rotations (|d rotvec| <= ~sqrt(2) |d R| entrywise near small angles)."""
    return np.repeat(relative_tolerance(g, parents, floor, k) * 1.5, 3)


def csr_of(g):
    """scipy CSR transfer matrix stored in a converter fixture (None: same topology)."""
    if 'csr_data' not in g:
        return None
    import scipy.sparse as sp

    n_out = len(g['csr_indptr']) - 1
    return sp.csr_matrix((g['csr_data'], g['csr_indices'], g['csr_indptr']),
                         shape=(n_out, int(g['csr_indices'].max()) + 1))
