import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def smpl_model():
    from human_dynamics_b200 import synthetic
    return synthetic.make_synthetic_smpl(seed=2)


@pytest.fixture(scope='session')
def smpl_model_dense():
    from human_dynamics_b200 import synthetic
    return synthetic.make_synthetic_smpl(seed=7, dense_weights=True, num_kps=19)


@pytest.fixture(scope='session')
def weights():
    from human_dynamics_b200 import synthetic
    return synthetic.make_synthetic_weights(seed=1, with_hal=True)
