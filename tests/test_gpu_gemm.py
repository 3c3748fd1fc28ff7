"""GPU: the pose-blend-shape contraction kernels in isolation -- tcgen05 (3xTF32 split, TMA,
TMEM) and the FP32 SIMT kernel against a float64 host product of the same operands."""

import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mname,Bp', [('smpl_tiny', 32), ('smpl_tiny', 160), ('smplx_tiny', 96), ('smpl', 256)])
def test_vposed_kernels(mname, Bp):
    from smplfitter_b200 import _native
    from smplfitter_b200.pt import BodyModel

    bm = BodyModel(mname).cuda()
    L = _native.lib()
    s = bm._struct()
    P = 9 * (bm.num_joints - 1)
    Kp = (P + 15) // 16 * 16
    rs = np.random.RandomState(0)
    feat = np.zeros((Bp, Kp), np.float32)
    feat[:, :P] = rs.randn(Bp, P).astype(np.float32)
    d_feat = torch.from_numpy(feat).cuda()
    rows = 3 * bm.num_vertices
    want = (bm._t_v_template_fit.double().cpu().numpy()[:, None]
            + bm._t_posedirs_fit.double().cpu().numpy()[:, :P] @ feat[:, :P].astype(np.float64).T)
    scale = np.abs(bm._t_posedirs_fit.cpu().numpy()).max() * np.sqrt(P)
    scratch = torch.empty(L.smplfit_debug_vposed_scratch_bytes(C.byref(s), Bp), dtype=torch.uint8, device='cuda')
    errs = {}
    for use_tc in (0, 1):
        out = torch.full((rows, Bp), float('nan'), device='cuda')
        _native.check(L.smplfit_debug_vposed(C.byref(s), d_feat.data_ptr(), Bp, use_tc, out.data_ptr(),
                                             scratch.data_ptr(), _native.stream_ptr(out.device)))
        torch.cuda.synchronize()
        errs[use_tc] = float(np.abs(out.cpu().numpy() - want).max())
    print(mname, Bp, 'simt err', errs[0], 'tc err', errs[1], 'scale', scale)
    # fp32-GEMM grade: a few ulps of the accumulated magnitude
    assert errs[0] < 2e-6 * max(1.0, scale * 10)
    assert errs[1] < 2e-6 * max(1.0, scale * 10)
