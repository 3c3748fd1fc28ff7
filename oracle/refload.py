"""TEST INFRASTRUCTURE -- loads the *unmodified* reference with the synthetic model injected
through its single data seam ``smplfitter.common.initialize`` (common.py:219; called from
pt/bodymodel.py:68 and np/bodymodel.py:47).

Search order: ``$SMPLFITTER_REFERENCE_SRC``, ``/root/reference/src`` (build container), then
``oracle/_ref`` (the offline install made by ``oracle/stage_ref.py``; this is what exists on the
GPU box).  Used by ``oracle/make_golden.py`` (pins the numpy restatement, writes
``tests/golden/*.npz``), by the ``-m gpu`` test that compares the CUDA path with the reference's
pt backend on the same B200, and by ``bench.py``'s reference legs.
"""

from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _candidates():
    env = os.environ.get('SMPLFITTER_REFERENCE_SRC')
    if env:
        yield env
    yield '/root/reference/src'
    yield os.path.join(HERE, '_ref')


def source_dir():
    for c in _candidates():
        if os.path.isfile(os.path.join(c, 'smplfitter', 'pt', 'bodyfitter.py')):
            return c
    return None


def available() -> bool:
    return source_dir() is not None


def load():
    """Return the reference ``smplfitter`` package with ``common.initialize`` patched."""
    src = source_dir()
    if src is None:
        raise RuntimeError('reference package not found (neither /root/reference/src nor oracle/_ref)')
    if src not in sys.path:
        sys.path.insert(0, src)
    import smplfitter.common as ref_common
    from smplfitter_b200 import modeldata

    modeldata.use_synthetic_models(True)
    ref_common.initialize = modeldata.initialize
    import smplfitter.pt  # noqa: F401  (binds smplfitter_common.initialize lazily via module attr)

    return sys.modules['smplfitter']
