#!/usr/bin/env python
"""Benchmark of the hot path: SMPL fits per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME] [--scaling weak|strong]

Default workload (BASELINE.json configs[1]): synthetic SMPL-shaped model (6890 vertices, 24 joints,
10 betas), BodyFitter.fit(num_iter=3, beta_regularizer=1, final_adjust_rots=True,
requested_keys=['pose_rotvecs','shape_betas']) on on-manifold targets, batch 4096 per GPU.
One "step" = one fit() call over the batch.  N > 1: one process per GPU (torchrun), the batch
dimension is sharded with no data-path collective (weak scaling: 4096 fits per rank per step;
``--scaling strong`` shards a global batch of 4096 instead).  ``--config`` selects the other BASELINE.json
configurations (smplx = configs[2], subset = configs[3], converter = configs[4]); they print the same
kind of line for their own workload and are not the headline.

`value`   : fits/s with the inputs already resident in HBM (CUDA events, max over ranks).
`e2e`     : the same through the public API from pinned HOST buffers, H2D/D2H copies inside the
            timed region.
`roofline`: the dominant kernel of the step, timed live with CUDA events on its stream in an
            instrumented repeat of the timed steps; algorithmic bytes per instance are stated in
            DESIGN.md.  Since round 2 the dominant kernel is k_fit_fused (blend-shape GEMM with the vertex pass in its
            epilogue): its algorithmic HBM bytes are the targets alone, and `roofline.tensor` adds the tensor-core rate.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference's own numpy backend
            (smplfitter.np.BodyFitter.fit from oracle/_ref, staged by build()) on the box's host cores on a
            bounded sample of the same workload; falls back to the numpy port (oracle/oracle_np.py,
            kind "port") only when the staged package is missing.
`reference_pt_b200`: the unmodified reference pt backend (eager and torch.jit.script) on the same GPU at the
            same batch size, and `v2v_mm_vs_reference` = mean vertex distance between the meshes re-posed from this
            library's fit and from the reference's fit of the same targets.
`scatter_gather` (N > 1): fits/s of smplfitter_b200.dist.scatter_fit_gather over NCCL -- the global batch lives on
            rank 0's GPU, is scattered, fitted and the results gathered back on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

E2E_CHUNK = 512
FIT_KW = dict(num_iter=3, beta_regularizer=1.0, final_adjust_rots=True,
              requested_keys=['pose_rotvecs', 'shape_betas'])

# name -> workload description (BASELINE.json configs[1..4])
CONFIGS = {
    'smpl': dict(model='smpl', mkw={}, batch=4096, kind='fit',
                 metric='SMPL fits/sec at batch 4096 (num_iter=3, 6890 verts, 24 joints, 10 betas)',
                 workload='SMPL fit num_iter=3 batch=4096 per GPU (BASELINE.json configs[1])',
                 model_desc='synthetic SMPL-shaped 6890v/24j/10b'),
    'smplx': dict(model='smplx', mkw={}, batch=4096, kind='fit',
                  metric='SMPL-X fits/sec at batch 4096 (num_iter=3, 10475 verts, 55 joints, 16 betas)',
                  workload='SMPL-X fit num_iter=3 batch=4096 per GPU (BASELINE.json configs[2])',
                  model_desc='synthetic SMPL-X-shaped 10475v/55j/16b'),
    'subset': dict(model='smpl', mkw=dict(vertex_subset_size=1024), batch=16384, kind='fit',
                   metric='SMPL 1024-vertex-subset fits/sec at batch 16384 per GPU (num_iter=3)',
                   workload='SMPL 1024-vertex subset fit num_iter=3 batch=16384 per GPU (BASELINE.json configs[3])',
                   model_desc='synthetic SMPL-shaped, 1024-vertex subset, 24j/10b'),
    'converter': dict(model='smpl', mkw={}, batch=4096, kind='convert',
                      metric='SMPL->SMPL-X conversions/sec at batch 4096 per GPU (forward LBS + transfer + fit, num_iter=1)',
                      workload='BodyConverter SMPL->SMPL-X batch=4096 per GPU (BASELINE.json configs[4])',
                      model_desc='synthetic SMPL 6890v/24j/10b -> synthetic SMPL-X 10475v/55j/16b, 3-nnz barycentric CSR'),
}


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms while the timed regions (resident and end-to-end) run."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20',
                 '-i', str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm)}


def synth_params(B, J, S, seed):
    rs = np.random.RandomState(seed)  # SURVEY.md 8d input distribution
    pose = (rs.randn(B, 3 * J) * 0.1).astype(np.float32)
    betas = (rs.randn(B, S) * 0.5).astype(np.float32)
    trans = rs.randn(B, 3).astype(np.float32)
    return pose, betas, trans


def synthetic_transfer_csr(t_in, t_out, seed=0):
    """Barycentric-style SMPL -> SMPL-X transfer: every output vertex from its 3 nearest input template vertices
    (SURVEY.md 8e': stands in for the licensed smpl2smplx_deftrafo_setup.pkl)."""
    import scipy.sparse as sp
    import scipy.spatial

    rs = np.random.RandomState(seed)
    _, cols = scipy.spatial.cKDTree(t_in).query(t_out, k=3)
    w = rs.dirichlet([2, 2, 2], size=len(t_out)).astype(np.float32)
    return sp.csr_matrix((w.reshape(-1), (np.repeat(np.arange(len(t_out)), 3), cols.reshape(-1))),
                         shape=(len(t_out), len(t_in)))


def algorithmic_bytes(kernel, B, V, J, S):
    """ALGORITHMIC HBM bytes of one launch of ``kernel`` over B instances (DESIGN.md section 4 states them per
    instance): what the kernel must read / write once, not what it happens to move."""
    v = B * 4 * 3 * V
    if 'k_fit_fused' in kernel:
        return v, ('fused blend-shape GEMM + vertex pass: the targets are read once (12 V bytes per instance), the posed '
                   'template stays in TMEM (average over the shape-stage and statistics launches)')
    if 'k_fwd_fused' in kernel:
        return v + B * 4 * (12 * J + S), 'forward LBS: (B,V,3) written once + joint rows / betas read'
    if 'k_vposed' in kernel:
        return v, 'pose-blend GEMM: v_posed^T written once (operands are L2-resident model constants + features)'
    if 'k_stats_tmpl' in kernel or 'k_mean' in kernel:
        return v, 'targets read once'
    if 'k_transpose' in kernel:
        return 2 * v, 'targets read once + instance-minor copy written once'
    if 'k_fwd_skin' in kernel:
        return 2 * v, 'v_posed^T read once + (B,V,3) written once'
    return 2 * v, 'vertex pass: targets + v_posed^T read once (2 x 12 V bytes per instance)'


def fused_tensor_rate(kernel, fitter, B, per_launch_ms):
    """Tensor-core work of one k_fit_fused launch: 2 x 128-instance tiles x padded rows x K flops per product, three fp16
    products per K step (hi*hi + hi*lo + lo*hi).  Reported next to the HBM figure; None for other kernels."""
    if 'k_fit_fused' not in kernel or fitter is None or getattr(fitter, '_fq', None) is None:
        return None
    rows = fitter._fq['fq_nseg_pad'] * 96
    bt = (B + 127) // 128 * 128
    useful = 2.0 * bt * rows * fitter._fq['fq_kf']
    peak = None
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peak = float(json.load(f)['bf16_tflops'])
    except Exception:
        pass
    issued = 3 * useful / (per_launch_ms / 1000) / 1e12
    return {'issued_tflops_fp16': issued, 'useful_tflops': issued / 3, 'peak_bf16_tflops': peak,
            'frac_issued': issued / peak if peak else None,
            'note': 'fp16-split GEMM: 3 products per K step give fp32-grade accuracy; issued = 3 x useful'}


def host_cores():
    """Threads a numpy implementation can actually use: the BLAS pool size (matmul / einsum calls are the only
    threaded part), capped by the scheduler affinity."""
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info

        blas = [p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') == 'blas']
        if blas:
            return int(min(avail, max(blas)))
    except Exception:
        pass
    return avail


class CpuBaseline:
    """The CPU implementation of the configured workload on the host cores: the unmodified reference's numpy
    backend when oracle/_ref is staged (kind 'reference'), else the numpy port (kind 'port')."""

    def __init__(self, cfg):
        from smplfitter_b200 import modeldata

        modeldata.use_synthetic_models(True)
        self.cfg = cfg
        self.kind = 'port'
        self.what = 'numpy port of smplfitter.pt (oracle/oracle_np.py); oracle/_ref is not staged'
        try:
            from oracle import refload

            if refload.available():
                refload.load()
                import smplfitter.np as rnp

                self.bm = rnp.BodyModel(cfg['model'], 'neutral', **cfg['mkw'])
                self.kind = 'reference'
                self.what = 'unmodified reference smplfitter.np (oracle/_ref)'
                if cfg['kind'] == 'convert':
                    self.bm_out = rnp.BodyModel('smplx', 'neutral')
                    self.fitter = rnp.BodyFitter(self.bm_out, enable_kid=True)
                    d_in, d_out = modeldata.initialize('smpl'), modeldata.initialize('smplx')
                    eye_in = np.tile(np.eye(3), [d_in.num_joints - 1, 1]).reshape(-1)
                    eye_out = np.tile(np.eye(3), [d_out.num_joints - 1, 1]).reshape(-1)
                    self.csr = synthetic_transfer_csr(d_in.v_template + d_in.posedirs @ eye_in,
                                                      d_out.v_template + d_out.posedirs @ eye_out)
                else:
                    self.fitter = rnp.BodyFitter(self.bm)
        except Exception as e:  # fall back to the port, say why
            self.kind = 'port'
            self.what = f'numpy port (oracle/oracle_np.py); reference import failed: {type(e).__name__}: {e}'
        if self.kind == 'port':
            from oracle import oracle_np

            self.om = oracle_np.OracleModel(modeldata.initialize(cfg['model'], **cfg['mkw']), cfg['model'])
            self.of = oracle_np.OracleFitter(self.om)
            if cfg['kind'] == 'convert':
                raise RuntimeError('the converter CPU baseline needs the staged reference (oracle/_ref)')

    def sample(self, B, seed):
        if self.kind == 'reference':
            J, S = self.bm.num_joints, self.bm.num_betas
        else:
            J, S = self.om.num_joints, self.om.num_betas
        return synth_params(B, J, S, seed)

    def step(self, params):
        """One pass of the workload over ``params``; returns seconds."""
        pose, betas, trans = params
        if self.cfg['kind'] == 'convert':
            t0 = time.perf_counter()
            v = self.bm(pose_rotvecs=pose, shape_betas=betas, trans=trans)['vertices']
            v = np.stack([self.csr @ v[i] for i in range(len(v))]).astype(np.float32)
            self.fitter.fit(v, num_iter=1, beta_regularizer=0.0, final_adjust_rots=False, kid_regularizer=1e9,
                            requested_keys=['pose_rotvecs', 'shape_betas'])
            return time.perf_counter() - t0
        if self.kind == 'reference':
            fw = self.bm(pose_rotvecs=pose, shape_betas=betas, trans=trans)
            tv, tj = np.ascontiguousarray(fw['vertices'], np.float32), np.ascontiguousarray(fw['joints'], np.float32)
            t0 = time.perf_counter()
            self.fitter.fit(tv, tj, **FIT_KW)
            return time.perf_counter() - t0
        fw = self.om.forward(pose, betas, trans)
        t0 = time.perf_counter()
        self.of.fit(fw['vertices'], fw['joints'], **FIT_KW)
        return time.perf_counter() - t0


def run_reference_arm(args, rank, cfg):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    base = CpuBaseline(cfg)
    unit = 'conversions/s' if cfg['kind'] == 'convert' else 'fits/s'
    sample_B = 32
    t_probe = base.step(base.sample(2, 7)) / 2  # warm caches / lazy imports; rough per-instance cost
    # bounded sample: keep the whole run (warm-up + steps) within ~2.5 minutes
    while sample_B > 4 and t_probe * sample_B * (args.warmup + args.steps) > 150:
        sample_B //= 2
    times = []
    for i in range(args.warmup + args.steps):
        dt = base.step(base.sample(sample_B, 42 + i))
        if i >= args.warmup:
            times.append(dt)
    total_t = sum(times)
    value = sample_B * len(times) / total_t
    line = {
        'impl': 'reference', 'metric': cfg['metric'], 'value': value, 'unit': unit, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000 * total_t / len(times),
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['workload'] + f'; each CPU step is a bounded sample of {sample_B} instances of it',
                   'model': cfg['model_desc']},
        'cpu_baseline': {'value': value, 'unit': unit, 'cores': host_cores(), 'kind': base.kind,
                         'sample': f'{sample_B} instances x {len(times)} steps; {base.what}; its hot ops are 3-operand '
                                   'einsums and a per-instance Python solve loop, so it effectively uses ~1 core'},
        'e2e': {'value': value, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def reference_pt_on_gpu(cfg, dev, tv, tj, fit_out, bm, B):
    """The unmodified reference pt backend on this GPU at the same batch (eager, then torch.jit.script), and the
    v2v distance between its fit and ours (``fit_out``) re-posed through the same forward."""
    import torch

    from oracle import refload

    if not refload.available():
        return None, None
    refload.load()
    import smplfitter.pt as rpt

    rbm = rpt.BodyModel(cfg['model'], 'neutral', **cfg['mkw']).to(dev)
    rfit = rpt.BodyFitter(rbm).to(dev)
    res = {}
    ref_out = None

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps, out

    try:
        ms, ref_out = timed(lambda: rfit.fit(tv, tj, **FIT_KW))
        res['eager_fits_per_s'] = B / ms * 1000
        res['eager_ms_per_step'] = ms
    except Exception as e:
        res['eager_error'] = f'{type(e).__name__}: {e}'[:200]
    try:
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            sfit = torch.jit.script(rfit)
        for _ in range(2):  # profiling executor warm-up
            sfit.fit(tv, tj, **FIT_KW)
        ms, _ = timed(lambda: sfit.fit(tv, tj, **FIT_KW))
        res['script_fits_per_s'] = B / ms * 1000
        res['script_ms_per_step'] = ms
    except Exception as e:
        res['script_error'] = f'{type(e).__name__}: {e}'[:200]
    res['batch'] = B
    res['what'] = 'unmodified smplfitter.pt.BodyFitter.fit (oracle/_ref) on the same GPU, same targets, CUDA events'
    v2v_mm = None
    if ref_out is not None:
        re_c = bm(fit_out['pose_rotvecs'], fit_out['shape_betas'], fit_out['trans'])['vertices']
        re_r = bm(ref_out['pose_rotvecs'], ref_out['shape_betas'], ref_out['trans'])['vertices']
        v2v_mm = float((re_c - re_r).norm(dim=-1).mean().item() * 1000)
        res['max_abs_diff'] = {k: float((fit_out[k] - ref_out[k]).abs().max().item())
                               for k in ('pose_rotvecs', 'shape_betas', 'trans')}
    del rfit, rbm
    torch.cuda.empty_cache()
    return res, v2v_mm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='smpl', choices=list(CONFIGS))
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--batch', type=int, default=None, help='instances per GPU per step (default: the config\'s)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-reference-gpu', action='store_true', help='skip the reference-pt-on-this-GPU leg')
    ap.add_argument('--resident-only', action='store_true',
                    help='profiling aid: only the resident timed steps (no e2e / forward / instrumented legs)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    cfg = CONFIGS[args.config]

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference_arm(args, rank, cfg)
        return

    import torch
    import torch.distributed as dist

    from smplfitter_b200 import _native, modeldata
    from smplfitter_b200 import dist as sdist
    from smplfitter_b200.pt import BodyConverter, BodyFitter, BodyModel

    modeldata.use_synthetic_models(True)
    sdist.bind_rank_to_local_cpus(local_rank)  # NUMA-local pinned buffers for the end-to-end leg
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    B_cfg = args.batch or cfg['batch']
    B = B_cfg if args.scaling == 'weak' else max(1, B_cfg // world)
    is_fit = cfg['kind'] == 'fit'
    unit = 'fits/s' if is_fit else 'conversions/s'
    bm = BodyModel(cfg['model'], **cfg['mkw']).to(dev)
    V, J, S = bm.num_vertices, bm.num_joints, bm.num_betas
    pose, betas, trans = synth_params(B, J, S, seed=42 + rank)
    d_pose, d_betas, d_trans = (torch.from_numpy(x).to(dev) for x in (pose, betas, trans))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if is_fit:
        fitter = BodyFitter(bm).to(dev)
        # synthetic on-manifold targets, a different shard per rank (inputs 340 MB > 126 MB L2)
        fw = bm(d_pose, d_betas, d_trans)
        tv, tj = fw['vertices'].contiguous(), fw['joints'].contiguous()
        h_tv = torch.empty(tv.shape, dtype=torch.float32, pin_memory=True).copy_(tv.cpu())
        h_tj = torch.empty(tj.shape, dtype=torch.float32, pin_memory=True).copy_(tj.cpu())
        h_out = {k: torch.empty(s, dtype=torch.float32, pin_memory=True)
                 for k, s in (('pose_rotvecs', (B, 3 * J)), ('shape_betas', (B, S)), ('trans', (B, 3)),
                              ('orientations', (B, J, 3, 3)), ('relative_orientations', (B, J, 3, 3)))}
        h2d = int(B * (V + J) * 12)
        d2h = int(sum(t.numel() for t in h_out.values()) * 4)

        def step_resident():
            return fitter.fit(tv, tj, **FIT_KW)

        def step_e2e():
            # public API for host-resident inputs (BodyFitter.fit_from_host -> smplfit_fit_host): chunked H2D
            # copies overlapped with the fits, every result key of fit() copied back into pinned host memory,
            # stream synchronised before it returns
            return fitter.fit_from_host(h_tv, h_tj, chunk_size=E2E_CHUNK, out=h_out, **FIT_KW)
    else:
        bm_out = BodyModel('smplx').to(dev)
        csr = synthetic_transfer_csr(bm._t_template_mesh.cpu().numpy(), bm_out._t_template_mesh.cpu().numpy())
        conv = BodyConverter(bm, bm_out, vertex_converter_csr=csr).to(dev)
        h_in = [torch.from_numpy(x).pin_memory() for x in (pose, betas, trans)]
        h_res = {k: torch.empty(s, dtype=torch.float32, pin_memory=True)
                 for k, s in (('pose_rotvecs', (B, 3 * bm_out.num_joints)), ('shape_betas', (B, bm_out.num_betas)),
                              ('trans', (B, 3)))}
        h2d = int(sum(t.numel() for t in h_in) * 4)
        d2h = int(sum(t.numel() for t in h_res.values()) * 4)

        def step_resident():
            return conv.convert(d_pose, d_betas, d_trans, num_iter=1)

        def step_e2e():
            p, b, t = (x.to(dev, non_blocking=True) for x in h_in)
            out = conv.convert(p, b, t, num_iter=1)
            for k, v in h_res.items():
                v.copy_(out[k], non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return h_res

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _native.launch_count(reset=True)
    ms = timed(step_resident, args.steps)
    launches = _native.launch_count()
    clocks = None

    if args.resident_only:
        if rank == 0:
            sampler.stop()
            _native.profile(True)
            torch.cuda.synchronize(dev)
            for _ in range(3):
                step_resident()
            torch.cuda.synchronize(dev)
            _native.profile(False)
            kprof = _native.profile_report()
            print(json.dumps({'resident_only': True, 'ms_per_step': ms / args.steps, 'gpu_launches': int(launches),
                              'kernel_ms_per_step': {k: round(v[1] / 3, 5) for k, v in sorted(kprof.items(), key=lambda kv: -kv[1][1])},
                              'note': 'profiling run, not a bench line'}))
        if world > 1:
            dist.destroy_process_group()
        return
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # the node's ceiling for the end-to-end leg: every rank copying its pinned targets to its GPU AT THE SAME TIME
    # (one host memory system / PCIe root feeds all GPUs), nothing else running
    ms_copy = None
    if is_fit:
        def step_copy():
            tv.copy_(h_tv, non_blocking=True)
            tj.copy_(h_tj, non_blocking=True)

        step_copy()
        ms_copy = timed(step_copy, 5) / 5

    # N > 1: the north_star's data movement -- the global batch on rank 0's GPU, NCCL scatter -> fit -> gather
    scatter_gather = None
    if world > 1 and is_fit:
        total = B * world
        if rank == 0:
            g_tv = tv.repeat(world, 1, 1)
            g_tj = tj.repeat(world, 1, 1)
        else:
            g_tv = g_tj = None

        def step_sg():
            return sdist.scatter_fit_gather(fitter.fit, total, g_tv, g_tj, V, J, device=dev, **FIT_KW)

        for _ in range(2):
            step_sg()
        sg_steps = max(3, args.steps // 2)
        ms_sg = timed(step_sg, sg_steps)
        scatter_gather = {'fits_per_s': total * sg_steps / (ms_sg / 1000), 'ms_per_step': ms_sg / sg_steps,
                          'global_batch': total, 'scatter_bytes_per_step': int(total * (V + J) * 12 * (world - 1) / world),
                          'note': 'smplfitter_b200.dist.scatter_fit_gather over NCCL: targets resident on rank 0, '
                                  'chunked scatter overlapped with the fits, packed results gathered on rank 0'}
        del g_tv, g_tj

    # instrumented repeat: per-kernel device time with CUDA events on the launching stream
    prof = {}
    if rank == 0:
        _native.profile(True)
        torch.cuda.synchronize(dev)
        for _ in range(min(args.steps, 5)):
            step_resident()
        torch.cuda.synchronize(dev)
        _native.profile(False)
        prof = _native.profile_report()
    if world > 1:
        dist.barrier()

    out = step_resident()
    # latency of one small fit (the launch chain is what bounds it): median of 20 calls at batch 32, device time
    lat32 = None
    if is_fit and rank == 0:
        tv32, tj32 = tv[:32].contiguous(), tj[:32].contiguous()
        for _ in range(3):
            fitter.fit(tv32, tj32, **FIT_KW)
        ts = []
        for _ in range(20):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            a0.record()
            fitter.fit(tv32, tj32, **FIT_KW)
            a1.record()
            torch.cuda.synchronize(dev)
            ts.append(a0.elapsed_time(a1))
        lat32 = sorted(ts)[len(ts) // 2]
    v2v_mm = ref_gpu = v2v_ref_mm = lbs = None
    if is_fit:
        # parity spot check inside the bench run (not timed): v2v of the re-posed fit against the targets
        re = bm(out['pose_rotvecs'], out['shape_betas'], out['trans'])
        v2v_mm = float((re['vertices'] - tv).norm(dim=-1).mean().item() * 1000)
        if rank == 0 and world == 1 and not args.no_reference_gpu:
            try:
                ref_gpu, v2v_ref_mm = reference_pt_on_gpu(cfg, dev, tv, tj, out, bm, B)
            except Exception as e:
                ref_gpu = {'error': f'{type(e).__name__}: {e}'[:300]}

    # secondary measurement: the forward LBS pass (store-bound; north_star's LBS roofline)
    if rank == 0:
        for _ in range(3):
            bm(d_pose, d_betas, d_trans)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(args.steps):
            bm(d_pose, d_betas, d_trans)
        e1.record()
        torch.cuda.synchronize(dev)
        t_fwd = e0.elapsed_time(e1) / args.steps / 1000
        peak_hbm, _src = peaks()
        fwd_bytes = B * (4 * 3 * V + 4 * (3 * J + S + 3))  # 83 020 B per forward (SMPL)
        lbs = {'forwards_per_s': B / t_fwd, 'ms_per_call': t_fwd * 1000, 'achieved_gbs': fwd_bytes / t_fwd / 1e9,
               'frac_of_hbm_peak': fwd_bytes / t_fwd / 1e9 / peak_hbm, 'bytes_per_forward': fwd_bytes // B}
    if world > 1:
        dist.barrier()

    if rank == 0:
        total_units = B * world * args.steps
        value = total_units / (ms / 1000)
        e2e_value = total_units / (ms_e2e / 1000)
        peak, peak_src = peaks()
        roof = None
        if prof:
            n_prof = min(args.steps, 5)
            tot = sum(v[1] for v in prof.values())
            dom = max(prof, key=lambda k: prof[k][1])
            n, t = prof[dom]
            per_launch_ms = t / n
            alg_bytes, alg_note = algorithmic_bytes(dom, B, V, J, S)
            achieved = alg_bytes / (per_launch_ms / 1000) / 1e9
            traffic = None
            try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu capture
                with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
                    tj_ = json.load(f)
                traffic = tj_.get(dom) if args.config == 'smpl' else None
            except Exception:
                pass
            roof = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                    'avg_launch_ms': per_launch_ms, 'share_of_step': t / tot, 'algorithmic_bytes': alg_bytes,
                    'note': alg_note + '; the fit is FP32-issue bound by design (SURVEY.md 8d), the HBM fraction is '
                                       'reported as required',
                    'tensor': fused_tensor_rate(dom, fitter if is_fit else None, B, per_launch_ms),
                    'kernel_ms_per_step': {k: v[1] / n_prof for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
                    'launches_per_step': {k: v[0] / n_prof for k, v in prof.items()}}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                base = CpuBaseline(cfg)
                base.step(base.sample(2, 7))
                nb = 16 if is_fit and args.config != 'smplx' else 8
                cdt = base.step(base.sample(nb, 42))
                cpu = {'value': nb / cdt, 'unit': unit, 'cores': host_cores(), 'kind': base.kind,
                       'sample': f'{nb} instances of the same workload in {cdt:.1f} s; {base.what}'}
            except Exception as e:
                cpu = {'value': None, 'unit': unit, 'cores': host_cores(), 'kind': 'unavailable',
                       'sample': f'{type(e).__name__}: {e}'[:200]}
        line = {
            'metric': cfg['metric'], 'value': value, 'unit': unit, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': value / 9481.0 if args.config == 'smpl' else None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': cfg['workload'] if args.scaling == 'weak'
                       else cfg['workload'].replace('per GPU', f'GLOBAL, sharded over {world} GPUs'),
                       'model': cfg['model_desc'], 'batch_per_gpu': B,
                       'l2': f'inputs ({B * (V + J) * 12 / 1e6:.0f} MB per rank) larger than L2 (126 MB)' if is_fit and B * V * 12 > 126e6
                       else 'working set is streamed once per step; outputs of one step are not inputs of the next',
                       'vs_baseline_note': 'published 9481 fits/s is the reference on an RTX 3090 (README.md:15)',
                       'parallelism': f'batch-sharded x{world}, no data-path collective'},
            'e2e': {'value': e2e_value, 'unit': unit, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / args.steps, 'chunk': E2E_CHUNK if is_fit else None,
                    'pcie_floor_ms': h2d / 54e9 * 1000,
                    'aggregate_h2d_gbs': h2d * world / (ms_e2e / args.steps / 1000) / 1e9,
                    'concurrent_copy_ms': ms_copy,
                    'concurrent_copy_gbs': (h2d * world / (ms_copy / 1000) / 1e9) if ms_copy else None,
                    'frac_of_copy_ceiling': (ms_copy / (ms_e2e / args.steps)) if ms_copy else None,
                    'note': ('BodyFitter.fit_from_host: pinned host targets -> pinned host results, stream synchronised '
                             'every step' if is_fit else 'pinned host parameters -> BodyConverter.convert -> pinned host results')
                            + '; pcie_floor_ms = H2D bytes / 54 GB/s (measured pinned H2D rate of one GPU of the box, '
                              'scripts/e2e_diag.py); aggregate_h2d_gbs = all ranks\' H2D bytes / step time; '
                              'concurrent_copy_* = the same pinned buffers copied by all ranks at once with nothing else '
                              'running (the node\'s measured ceiling for this leg), frac_of_copy_ceiling = that time / e2e time'},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu,
            'v2v_mm_roundtrip': v2v_mm, 'v2v_mm_vs_reference': v2v_ref_mm, 'reference_pt_b200': ref_gpu,
            'lbs_forward': lbs, 'scatter_gather': scatter_gather, 'latency_batch32_ms': lat32,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
