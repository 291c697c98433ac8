"""pytest configuration: registers the `gpu` marker and puts the repo root and oracle/ on sys.path.

`-m "not gpu"` (CPU container): oracle vs golden vectors, host logic, C-ABI symbol check, gloo world-size-2 paths.
`-m gpu` (B200 box): parity of the CUDA path against the oracle and the golden vectors, through the C-ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
