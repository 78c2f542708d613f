import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "mgpu: needs TWO CUDA devices and NCCL (run with -m mgpu on a 2-GPU lease); the same code "
                            "path runs on one GPU through the loopback communicator in test_ranks_as_contexts_of_one_process")


def pytest_collection_modifyitems(config, items):
    # GPU tests must FAIL, not skip, when selected on a box whose GPU path is broken; they are
    # only deselected by the driver's -m "not gpu" on the CPU box.
    pass


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
