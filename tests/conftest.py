import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "id-grec_b200")
GOLDEN = os.path.join(REPO, "tests", "golden")
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_tiny():
    return np.load(os.path.join(GOLDEN, "tiny.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_quirks():
    return np.load(os.path.join(GOLDEN, "quirks.npz"), allow_pickle=False)


def _dataset_dir(tmp_root, name):
    import shutil
    d = os.path.join(tmp_root, name)
    os.makedirs(d, exist_ok=True)
    shutil.copy(os.path.join(GOLDEN, name + "_train.txt"), os.path.join(d, "train.txt"))
    shutil.copy(os.path.join(GOLDEN, name + "_test.txt"), os.path.join(d, "test.txt"))
    return d


@pytest.fixture(scope="session")
def golden_dirs(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("golden_ds"))
    return {n: _dataset_dir(root, n) for n in ("tiny", "quirks")}
