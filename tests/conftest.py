import os
import sys

import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
# the oracle is test infrastructure: only tests (and smoke / bench's CPU legs) import it
ORACLE_DIR = os.path.join(REPO, "oracle")
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_case_names():
    d = os.path.join(GOLDEN, "cases")
    return sorted(f[:-4] for f in os.listdir(d) if f.endswith(".npz") and f != "full_body_perturbed.npz")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
