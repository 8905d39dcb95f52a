import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_cases():
    with open(os.path.join(GOLDEN_DIR, "cases.json")) as f:
        return json.load(f)["cases"]


_outputs = None


def golden_output(case_id):
    global _outputs
    if _outputs is None:
        _outputs = np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz"))
    return _outputs[case_id]


@pytest.fixture(scope="session")
def make_input():
    from oracle.make_golden import make_input as mk
    return mk


def tolerance(mat):
    """|gpu - ref| <= 1e-5 * max(1, max|mat|)."""
    return 1e-5 * max(1.0, float(np.max(np.abs(mat))))
