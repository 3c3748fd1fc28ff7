"""Static index structures of the fitter (host side, numpy, device-agnostic).

These are the "vertex-partition index masks" that must be bit-exact with the reference's
``BodyFitter.__init__`` (/root/reference/src/smplfitter/pt/bodyfitter.py:36-233):
dominant-joint vertex partition with the SMPL toe->foot merge (:36-47), part buckets by
joint count (:81-97), adjustable parts (:101-104), used vertices (:109-114), the assembly
permutation (:149-156), kinematic-tree levels (:181-192) and final-adjust levels (:219-233).
On top of those, ``FitPlan`` derives the device-side layouts the CUDA kernels consume
(sparse skinning table, per-part segment lists of the statistics pass).
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SMPL_ADJUSTABLE = [1, 2, 4, 5, 7, 8, 16, 17, 18, 19]


@dataclass
class FitPlan:
    num_joints: int
    num_vertices: int
    is_smpl_family: bool
    parents: np.ndarray  # (J,) int32, parents[0] = -1
    part_assignment: np.ndarray  # (V,) int64
    part_vertex_selectors: list  # J arrays of int64
    children_and_self: list  # J lists
    descendants_and_self: list
    multi_joint_parts: list
    bone_parts: list
    leaf_parts: list
    adjustable_parts: list
    stat_parts: list
    used_vertex_indices: np.ndarray  # (N_used,) int64
    part_matrix: np.ndarray  # (J, N_used) float32 one-hot
    part_counts: np.ndarray  # (J,) float32
    center_matrix: np.ndarray  # (J, J) float32
    mjp_joint_membership: np.ndarray  # (n_multi, J) float32
    bone_pairs: np.ndarray  # (n_bone, 2) int64
    assemble_indices: np.ndarray  # (J,) int64
    fk_js: np.ndarray
    fk_ps: np.ndarray
    fk_level_sizes: list
    adj_parts: np.ndarray
    adj_level_sizes: list
    adj_last_level: int
    leveladj_supported: bool
    adj_n_joints: int
    adj_part_joints: np.ndarray
    cas_starts: list
    cas_flat: np.ndarray
    # device-side descriptors
    part_kind: np.ndarray = field(default=None)  # (J,) int32: 0 skip, 1 multi, 2 bone, 3 leaf, 4 copy
    part_copy_src: np.ndarray = field(default=None)  # (J,) int32 (toe -> foot), else self
    part_is_stat: np.ndarray = field(default=None)  # (J,) uint8
    part_is_adjustable: np.ndarray = field(default=None)  # (J,) uint8
    cas_table: np.ndarray = field(default=None)  # (J, max_cas) int32, -1 padded
    cas_count: np.ndarray = field(default=None)  # (J,) int32
    vertex_stat_part: np.ndarray = field(default=None)  # (V,) int32: part id if used else -1


def build_fit_plan(weights: np.ndarray, kintree_parents, model_name: str) -> FitPlan:
    """Derive every static index structure from skin weights + kinematic tree.

    ``weights`` must be the float32 skin weights the model holds (the reference arg-maxes
    the float32 buffer, pt/bodyfitter.py:36); ties resolve to the lowest joint index.
    """
    w = np.asarray(weights, dtype=np.float32)
    V, J = w.shape
    parents = [int(p) for p in kintree_parents]
    smpl_family = model_name.startswith('smpl')

    part = np.argmax(w, axis=1).astype(np.int64)
    if smpl_family:
        part[part == 10] = 7
        part[part == 11] = 8
    selectors = [np.nonzero(part == i)[0].astype(np.int64) for i in range(J)]

    cas = [[i] for i in range(J)]
    for i in range(1, J):
        cas[parents[i]].append(i)
    desc = [[i] for i in range(J)]
    for i in range(J - 1, 0, -1):
        desc[parents[i]].extend(desc[i])

    multi, bone, leaf = [], [], []
    for i in range(J):
        if smpl_family and i in (10, 11):
            continue
        n = len(cas[i])
        (multi if n >= 3 else bone if n == 2 else leaf).append(i)
    adjustable = list(SMPL_ADJUSTABLE) if smpl_family else list(range(J))
    stat_parts = sorted(set(bone + leaf + adjustable))

    used_mask = np.zeros(V, dtype=bool)
    for i in stat_parts:
        used_mask[selectors[i]] = True
    used = np.nonzero(used_mask)[0].astype(np.int64)
    part_matrix = np.zeros((J, len(used)), dtype=np.float32)
    part_matrix[part[used], np.arange(len(used))] = 1.0
    part_counts = part_matrix.sum(axis=1)

    center = np.zeros((J, J), dtype=np.float32)
    for i in range(J):
        center[i, cas[i]] = np.float32(1.0) / np.float32(len(cas[i]))
    mjp = np.zeros((len(multi), J), dtype=np.float32)
    for k, i in enumerate(multi):
        mjp[k, cas[i]] = 1.0
    bone_pairs = np.array([[cas[i][0], cas[i][1]] for i in bone], dtype=np.int64).reshape(-1, 2)

    order = multi + leaf + bone
    inv = [0] * J
    for pos, j in enumerate(order):
        inv[j] = pos
    if smpl_family:
        inv[10] = inv[7]
        inv[11] = inv[8]
    assemble = np.array(inv, dtype=np.int64)

    depth = [0] * J
    for i in range(1, J):
        depth[i] = depth[parents[i]] + 1
    levels = [[i for i in range(J) if depth[i] == d] for d in range(1, max(depth) + 1)]
    fk_js = [i for js in levels for i in js]
    fk_ps = [parents[i] for i in fk_js]

    adj_set = set(adjustable)
    joint_counts = {len(cas[i]) for i in adj_set}
    leveladj = smpl_family and len(joint_counts) == 1
    adj_levels = [[i for i in js if i in adj_set] for js in levels]
    adj_flat = [i for a in adj_levels for i in a]
    adj_last = max((k for k, a in enumerate(adj_levels) if a), default=-1)
    if leveladj:
        adj_n = next(iter(joint_counts))
        adj_joints = [j for i in adj_flat for j in cas[i]]
    else:
        adj_n, adj_joints = 0, []
    cas_starts = [0]
    for i in range(J):
        cas_starts.append(cas_starts[-1] + len(cas[i]))

    plan = FitPlan(
        num_joints=J, num_vertices=V, is_smpl_family=smpl_family,
        parents=np.array([-1] + parents[1:], dtype=np.int32),
        part_assignment=part, part_vertex_selectors=selectors,
        children_and_self=cas, descendants_and_self=desc,
        multi_joint_parts=multi, bone_parts=bone, leaf_parts=leaf,
        adjustable_parts=adjustable, stat_parts=stat_parts,
        used_vertex_indices=used, part_matrix=part_matrix, part_counts=part_counts,
        center_matrix=center, mjp_joint_membership=mjp, bone_pairs=bone_pairs,
        assemble_indices=assemble,
        fk_js=np.array(fk_js, dtype=np.int64), fk_ps=np.array(fk_ps, dtype=np.int64),
        fk_level_sizes=[len(js) for js in levels],
        adj_parts=np.array(adj_flat, dtype=np.int64),
        adj_level_sizes=[len(a) for a in adj_levels], adj_last_level=adj_last,
        leveladj_supported=leveladj, adj_n_joints=adj_n,
        adj_part_joints=np.array(adj_joints, dtype=np.int64),
        cas_starts=cas_starts,
        cas_flat=np.array([j for js in cas for j in js], dtype=np.int64),
    )

    kind = np.zeros(J, dtype=np.int32)
    kind[multi] = 1
    kind[bone] = 2
    kind[leaf] = 3
    src = np.arange(J, dtype=np.int32)
    if smpl_family:
        kind[10] = 4
        kind[11] = 4
        src[10] = 7
        src[11] = 8
    plan.part_kind = kind
    plan.part_copy_src = src
    is_stat = np.zeros(J, dtype=np.uint8)
    is_stat[stat_parts] = 1
    plan.part_is_stat = is_stat
    is_adj = np.zeros(J, dtype=np.uint8)
    is_adj[adjustable] = 1
    plan.part_is_adjustable = is_adj
    max_cas = max(len(c) for c in cas)
    table = -np.ones((J, max_cas), dtype=np.int32)
    for i in range(J):
        table[i, : len(cas[i])] = cas[i]
    plan.cas_table = table
    plan.cas_count = np.array([len(c) for c in cas], dtype=np.int32)
    vsp = np.where(is_stat[part] > 0, part, -1).astype(np.int32)
    plan.vertex_stat_part = vsp
    return plan


def sparse_skin_table(weights: np.ndarray, max_k: int = 8):
    """Top-K (index, weight) table of the skin weights.

    Returns ``(idx (V,K) int32, w (V,K) float32, K)`` with K = max non-zeros per row; rows
    are padded with weight 0 on the row's dominant joint.  Models whose rows hold more than
    ``max_k`` non-zeros fall back to the dense table (K = J).
    """
    w = np.asarray(weights, dtype=np.float32)
    V, J = w.shape
    nnz = (w != 0).sum(axis=1)
    K = int(nnz.max())
    if K > max_k:
        K = J
    order = np.argsort(-(w != 0).astype(np.int8), axis=1, kind='stable')[:, :K]  # non-zeros first, by joint index
    idx = order.astype(np.int32)
    ww = np.take_along_axis(w, order, axis=1).astype(np.float32)
    dom = np.argmax(w, axis=1).astype(np.int32)
    pad = ww == 0
    idx = np.where(pad, dom[:, None], idx).astype(np.int32)
    return idx, ww, K
