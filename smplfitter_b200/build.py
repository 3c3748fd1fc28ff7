"""Builds libsmplfit_b200.so in-tree with nvcc for sm_100a (no GPU needed: cross-compiles).

    python -m smplfitter_b200.build [--force]
"""

from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
LIB = os.path.join(HERE, 'libsmplfit_b200.so')
SOURCES = ['fit.cu', 'fit_host.cu', 'pass_shape.cu', 'pass_shape_v3.cu', 'pass_lite.cu', 'pass_stats.cu', 'pass_scale.cu', 'forward.cu', 'fwd_fused.cu', 'fit_fused.cu', 'vposed_tc.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '-DSMPLFIT_BUILD',
    '-Xfatbin=-compress-all',  # (file size only: the device code is compressed in the fat binary, decompressed at load)
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _deps(path: str, seen: set) -> None:
    """The file and, recursively, every local header it includes."""
    if path in seen or not os.path.exists(path):
        return
    seen.add(path)
    with open(path, 'r', errors='replace') as f:
        for line in f:
            if line.lstrip().startswith('#include "'):
                inc = line.split('"')[1]
                _deps(os.path.normpath(os.path.join(os.path.dirname(path), inc)), seen)


def _digest(src: str) -> str:
    seen: set = set()
    _deps(os.path.join(CSRC, src), seen)
    h = hashlib.sha256()
    for path in sorted(seen):
        with open(path, 'rb') as f:
            h.update(os.path.basename(path).encode())
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True, out: str | None = None, extra: list | None = None) -> str:
    """Default: the in-tree library, one object per translation unit, each rebuilt only when the source or a header it
    includes changed.  ``out`` / ``extra`` build a variant (extra nvcc flags, e.g. a -D tuning macro) into another
    file for A/B runs (load it with SMPLFIT_B200_LIB=...)."""
    if out or extra:
        return _build_variant(out or LIB, list(extra or []), verbose)
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        stamp, digest = obj + '.stamp', _digest(src)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
            return obj, False
        cmd = [nvcc, *NVCC_FLAGS, '-I', INCLUDE, '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        if r.stderr.strip() and verbose:
            print(r.stderr.strip())
        with open(stamp, 'w') as f:
            f.write(digest)
        return obj, True

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in res]
    if not any(c for _, c in res) and os.path.exists(LIB) and not force:
        return LIB
    cmd = [nvcc, '-shared', '-o', LIB, *objs, '-lcuda']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    if verbose:
        print(f'built {LIB} ({sum(c for _, c in res)} of {len(res)} objects recompiled)')
    return LIB


def _build_variant(out: str, extra: list, verbose: bool) -> str:
    tag = hashlib.sha256((' '.join(extra) + out).encode()).hexdigest()[:10]
    objdir = os.path.join(HERE, 'build', 'variant_' + tag)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        r = subprocess.run([nvcc, *NVCC_FLAGS, *extra, '-I', INCLUDE, '-c', os.path.join(CSRC, src), '-o', obj],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{r.stdout}\n{r.stderr}')
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, '-shared', '-o', out, *objs, '-lcuda'], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    if verbose:
        print(f'built {out} with {extra}')
    return out


if __name__ == '__main__':
    _out = sys.argv[sys.argv.index('--out') + 1] if '--out' in sys.argv else None
    _extra = sys.argv[sys.argv.index('--extra') + 1].split() if '--extra' in sys.argv else None
    build(force='--force' in sys.argv, out=_out, extra=_extra)
