import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_case(name):
    """Golden case (tests/golden/<name>.npz) as a dict of numpy arrays, plus the parsed spec params."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = {k: z[k] for k in z.files}
    with open(os.path.join(GOLDEN, "specs", str(case["spec"]) + ".json")) as f:
        case["params"] = json.load(f)["params"]
    return case


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("dataset_") and "_train" not in f)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
