import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["tiny", "small", "small_sparse_ids", "small_ties", "small_float", "edge_mix"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_inputs(g):
    return json.loads(str(g["in_gt_json"])), json.loads(str(g["in_dt_json"]))


@pytest.fixture(scope="session", params=GOLDEN_CASES)
def golden(request):
    g = load_golden(request.param)
    g["_name"] = request.param
    return g
