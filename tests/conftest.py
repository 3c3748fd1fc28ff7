import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

# the licensed SMPL files are absent: every test runs on the synthetic stand-ins (explicit opt-in;
# without it BodyModel('smpl') raises FileNotFoundError like the reference)
from smplfitter_b200 import modeldata  # noqa: E402

modeldata.use_synthetic_models(True)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN_DIR
