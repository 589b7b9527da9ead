import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.build()
    return o


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cnc_b200 import _lib

    _lib.lib()  # fail loudly if the extension is missing on a GPU box
    return torch.device("cuda:0")


# product layout (train_CNC_nerf_synthetic.py:150-155)
R3 = [18, 24, 33, 44, 59, 80, 108, 148, 201, 275, 376, 514]
R2 = [130, 258, 514, 1026]
R16 = [int(np.floor(16 * (512 / 16) ** (l / 15))) + 2 for l in range(16)]
