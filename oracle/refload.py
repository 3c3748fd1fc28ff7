"""TEST INFRASTRUCTURE -- loads the *unmodified* reference (``/root/reference/src``) with the
synthetic model injected through its single data seam ``smplfitter.common.initialize``
(common.py:219; called from pt/bodymodel.py:68).  Only usable in the build container (the
reference tree does not exist on the GPU box); used by ``oracle/make_golden.py`` to pin the
numpy restatement and to generate ``tests/golden/*.npz``.
"""

from __future__ import annotations

import os
import sys

REF_SRC = os.environ.get('SMPLFITTER_REFERENCE_SRC', '/root/reference/src')


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, 'smplfitter'))


def load():
    """Return the reference ``smplfitter`` package with ``common.initialize`` patched."""
    if not available():
        raise RuntimeError(f'reference sources not found at {REF_SRC}')
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import smplfitter.common as ref_common
    from smplfitter_b200 import modeldata

    ref_common.initialize = modeldata.initialize
    import smplfitter.pt  # noqa: F401  (binds smplfitter_common.initialize lazily via module attr)

    return sys.modules['smplfitter']
