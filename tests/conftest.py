import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def state_dict():
    """The seeded weights every fixture in tests/golden was generated with (oracle/gen_golden.py)."""
    from founddiff_b200 import weights
    return weights.random_state_dict(10)


@pytest.fixture(scope="session")
def state_dict1():
    """Weights of the second Unet (num_unet = 2) in tests/golden/objectives.npz (oracle/gen_golden_objectives.py)."""
    from founddiff_b200 import weights
    return weights.random_state_dict(11)


# tag -> (num_unet, objective, test_res_or_noise) of tests/golden/objectives.npz
OBJECTIVE_CONFIGS = dict(rn=(2, "pred_res_noise", "res_noise"), rn_res=(2, "pred_res_noise", "res"),
                         rn_noise=(2, "pred_res_noise", "noise"), x0n=(2, "pred_x0_noise", "res_noise"),
                         noise=(1, "pred_noise", "None"))
