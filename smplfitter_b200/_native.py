"""ctypes binding of libsmplfit_b200.so (the C ABI in include/smplfit_b200.h).

The library is the product: if it is missing or a call fails this module raises -- there is
no CPU or eager-PyTorch fallback anywhere in ``smplfitter_b200``.
"""

from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

_F = C.c_void_p  # device pointers travel as plain addresses


class ModelStruct(C.Structure):
    """Mirror of ``smplfit_model_t`` (field order must match the header)."""

    _int_fields = [
        'num_vertices', 'num_joints', 'num_betas', 'num_pose_feats', 'skin_k', 'is_smpl_family',
        'n_used', 'n_segments', 'chunk_len', 'max_cas', 'n_adjustable', 'reserved1',
    ]
    _ptr_fields_a = [
        'v_template', 'shapedirs', 'posedirs', 'kid_shapedir', 'J_template', 'J_shapedirs',
        'kid_J_shapedir', 'J_regressor', 'template_mesh', 'parents', 'skin_idx', 'skin_w', 'order',
        'seg_start', 'seg_part', 'part_seg_begin', 'part_kind', 'part_copy_src', 'part_flags',
        'cas_table', 'cas_count', 'inv_order', 'posedirs_fit', 'v_template_fit',
    ]
    _ptr_fields_b = [
        'fit_shapedirs', 'fit_Jt_ext', 'template_joints_regressed', 'J_regressor_fit', 'posedirs_hi',
        'posedirs_lo', 'template_mesh_fit', 'fit_rec', 'fit_wS', 'fit_wsum', 'fwd_rec',
    ]
    _fields_ = (
        [(n, C.c_int32) for n in _int_fields]
        + [(n, _F) for n in _ptr_fields_a]
        + [('fit_ns', C.c_int32), ('fit_reserved', C.c_int32)]
        + [(n, _F) for n in _ptr_fields_b]
        + [('fit_rec_len', C.c_int32), ('fwd_rec_len', C.c_int32)]
        + [(n, _F) for n in ('seg_slots', 'yj_start', 'yj_entry', 'gcf_pairs', 'gcf_A', 'gcf_G0', 'gcf_lstart',
                             'gcf_lk', 'gcf_Bm', 'gcf_Wh')]
        + [('n_slots', C.c_int32), ('gcf_npairs', C.c_int32)]
        + [('gcf_AT_hi', _F), ('gcf_AT_lo', _F), ('posedirs_model_hi', _F), ('posedirs_model_lo', _F), ('posedirs_model_f32', _F), ('fit_slot_mask', _F)]
        + [('fwd_P_hi', _F), ('fwd_P_lo', _F), ('fwd_vrec', _F), ('fwd_kf', C.c_int32), ('fwd_scale_log2', C.c_int32)]
        + [('fit_P_hi', _F), ('fit_P_lo', _F), ('fit_kf', C.c_int32), ('fit_scale_log2', C.c_int32)]
        + [('fq_P_hi', _F), ('fq_P_lo', _F), ('fq_rec', _F), ('fq_sd', _F), ('fq_kf', C.c_int32),
           ('fq_scale_log2', C.c_int32), ('fq_sdl', C.c_int32), ('fq_nseg_pad', C.c_int32)]
        + [('jreg_ptr', _F), ('jreg_idx', _F), ('jreg_val', _F)]
    )


class FitOpts(C.Structure):
    """Mirror of ``smplfit_fit_opts_t``."""

    _fields_ = [
        ('num_iter', C.c_int32), ('final_adjust_rots', C.c_int32), ('enable_kid', C.c_int32),
        ('want_pose_rotvecs', C.c_int32), ('want_rel_orient', C.c_int32), ('shape_weights', C.c_int32),
        ('scale_mode', C.c_int32), ('share_beta', C.c_int32),
        ('beta_regularizer', C.c_float), ('beta_regularizer2', C.c_float),
        ('kid_regularizer', C.c_float), ('scale_regularizer', C.c_float),
    ]


ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)


def lib_path() -> str:
    return os.environ.get('SMPLFIT_B200_LIB', os.path.join(_HERE, 'libsmplfit_b200.so'))


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f'{path} not found: build the CUDA library first (python -m smplfitter_b200.build). '
            'smplfitter_b200 has no fallback path.'
        )
    L = C.CDLL(path)
    L.smplfit_version.restype = C.c_char_p
    L.smplfit_last_error.restype = C.c_char_p
    L.smplfit_struct_size.restype = C.c_size_t
    L.smplfit_struct_size.argtypes = [C.c_int]
    L.smplfit_launch_count.restype = C.c_int64
    L.smplfit_launch_count.argtypes = [C.c_int]
    L.smplfit_graph_replays.restype = C.c_int64
    L.smplfit_graph_replays.argtypes = [C.c_int]
    L.smplfit_graph_stats.restype = None
    L.smplfit_graph_stats.argtypes = [C.POINTER(C.c_int64)]
    L.smplfit_forward_workspace_bytes.restype = C.c_size_t
    L.smplfit_forward_workspace_bytes.argtypes = [C.POINTER(ModelStruct), C.c_int64]
    L.smplfit_forward.restype = C.c_int
    L.smplfit_forward.argtypes = [
        C.POINTER(ModelStruct), C.c_int64, C.c_int, _F, _F, C.c_int, _F, _F, _F, _F, _F, _F, C.c_size_t, _F,
    ]
    L.smplfit_fit_workspace_bytes.restype = C.c_size_t
    L.smplfit_fit_workspace_bytes.argtypes = [
        C.POINTER(ModelStruct), C.c_int64, C.POINTER(FitOpts), C.c_int, C.c_int, C.c_int,
    ]
    L.smplfit_fit.restype = C.c_int
    L.smplfit_fit.argtypes = (
        [C.POINTER(ModelStruct), C.c_int64] + [_F] * 9 + [C.POINTER(FitOpts)] + [_F] * 7
        + [_F, C.c_size_t, _F]
    )
    L.smplfit_fit_host_workspace_bytes.restype = C.c_size_t
    L.smplfit_fit_host_workspace_bytes.argtypes = [
        C.POINTER(ModelStruct), C.c_int64, C.c_int64, C.POINTER(FitOpts), C.c_int,
    ]
    L.smplfit_fit_host.restype = C.c_int
    L.smplfit_fit_host.argtypes = (
        [C.POINTER(ModelStruct), C.c_int64, C.c_int64, _F, _F, C.POINTER(FitOpts)] + [_F] * 7
        + [_F, C.c_size_t, _F]
    )
    L.smplfit_fit_known_pose.restype = C.c_int
    L.smplfit_fit_known_pose.argtypes = (
        [C.POINTER(ModelStruct), C.c_int64] + [_F] * 7 + [C.POINTER(FitOpts)] + [_F] * 5
        + [_F, C.c_size_t, _F]
    )
    L.smplfit_profile.restype = C.c_int
    L.smplfit_profile.argtypes = [C.c_int]
    L.smplfit_profile_report.restype = C.c_int
    L.smplfit_profile_report.argtypes = [C.c_char_p, C.c_size_t]
    L.smplfit_debug_vposed_scratch_bytes.restype = C.c_size_t
    L.smplfit_debug_vposed_scratch_bytes.argtypes = [C.POINTER(ModelStruct), C.c_int]
    L.smplfit_debug_vposed.restype = C.c_int
    L.smplfit_debug_vposed.argtypes = [C.POINTER(ModelStruct), _F, C.c_int, C.c_int, _F, _F, _F]
    L.smplfit_fit_known_shape.restype = C.c_int
    L.smplfit_fit_known_shape.argtypes = (
        [C.POINTER(ModelStruct), C.c_int64, _F, C.c_int] + [_F] * 8 + [C.POINTER(FitOpts)] + [_F] * 5
        + [_F, C.c_size_t, _F]
    )
    L.smplfit_set_share_beta_allreduce.restype = C.c_int
    L.smplfit_set_share_beta_allreduce.argtypes = [ALLREDUCE_FN, C.c_void_p, C.c_int64]
    L.smplfit_convert_vertices.restype = C.c_int
    L.smplfit_convert_vertices.argtypes = [_F, _F, _F, C.c_int32, C.c_int32, C.c_int64, _F, _F, _F]
    if L.smplfit_struct_size(0) != C.sizeof(ModelStruct) or L.smplfit_struct_size(1) != C.sizeof(FitOpts):
        raise RuntimeError('smplfit_b200: struct layout mismatch between _native.py and the library')
    _lib = L
    return L


def check(code: int) -> None:
    if code != 0:
        msg = lib().smplfit_last_error().decode()
        if code == -2:
            raise NotImplementedError(f'smplfit_b200: {msg}')
        raise RuntimeError(f'smplfit_b200 error {code}: {msg}')


def ptr(t) -> int:
    """Device address of a tensor (0 for None)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count(reset: bool = False) -> int:
    return int(lib().smplfit_launch_count(1 if reset else 0))


def graph_replays(reset: bool = False) -> int:
    """Fits served by replaying a captured CUDA graph since the last reset."""
    return int(lib().smplfit_graph_replays(1 if reset else 0))


def graph_stats() -> dict:
    buf = (C.c_int64 * 4)()
    lib().smplfit_graph_stats(buf)
    return dict(zip(('first_sights', 'captures', 'capture_failures', 'instantiated'), [int(x) for x in buf]))


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f'smplfitter_b200: {what} lives on {t.device}; this package only runs on a CUDA device '
            '(move the module and its inputs with .cuda()). There is no CPU fallback.'
        )


def profile(enable: bool) -> None:
    lib().smplfit_profile(1 if enable else 0)


def profile_report() -> dict:
    """{kernel name: (launches, total_ms)} of the launches made while profiling was enabled."""
    buf = C.create_string_buffer(1 << 16)
    lib().smplfit_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split('\t')
        name = name.strip('()').split('<')[0]
        cnt, tot = out.get(name, (0, 0.0))
        out[name] = (cnt + int(n), tot + float(ms))
    return out
