"""TEST INFRASTRUCTURE -- stages the UNMODIFIED reference package into ``oracle/_ref``.

The reference (``/root/reference``, pure Python) does not exist on the GPU box, so the build
container installs it once into the git-ignored (but not gpurun-ignored) directory
``oracle/_ref`` with pip's offline ``--target`` install.  Nothing of it enters the repository
history; on the GPU box the tests / ``bench.py --impl reference`` import it from there
(``oracle/refload.py``), with the synthetic model injected through the reference's single data
seam ``smplfitter.common.initialize``.

    python -m oracle.stage_ref           # (re)install when /root/reference is present

Called by ``__graft_entry__.build()``.  The source tree is read-only and the build writes an
egg-info directory, so the install runs from a copy under /tmp.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('SMPLFITTER_REFERENCE_ROOT', '/root/reference')
TARGET = os.path.join(HERE, '_ref')


def staged() -> bool:
    return os.path.isfile(os.path.join(TARGET, 'smplfitter', 'pt', 'bodyfitter.py'))


def stage(force: bool = False, verbose: bool = True) -> bool:
    """Returns True when ``oracle/_ref`` holds the reference package afterwards."""
    if staged() and not force:
        return True
    if not os.path.isdir(os.path.join(REF_ROOT, 'src', 'smplfitter')):
        return staged()
    tmp = tempfile.mkdtemp(prefix='smplfitter_ref_')
    try:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(REF_ROOT, src, ignore=shutil.ignore_patterns('.git', '__pycache__'))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        env = dict(os.environ, SETUPTOOLS_SCM_PRETEND_VERSION='0.0.0')
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps',
               '--find-links', '/opt/wheelhouse', '--target', TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0 or not staged():
            raise RuntimeError(f'offline install of the reference failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}')
        if verbose:
            print(f'staged the unmodified reference into {TARGET}')
        return True
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == '__main__':
    ok = stage(force='--force' in sys.argv)
    print('staged' if ok else 'reference tree not available; nothing staged')
