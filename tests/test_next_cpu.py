"""CPU checks for the section-8(f) rows: activity split of the test users, the oracle restatements
of the added losses against outputs of the unmodified reference (tests/golden/next.npz, produced by
tests/golden/make_golden_next.py), and the closed-form gradients the CUDA kernels implement."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def golden_next():
    return np.load(os.path.join(GOLDEN, "next.npz"), allow_pickle=False)


@pytest.mark.parametrize("name", ["tiny", "quirks"])
def test_sparsity_split_matches_reference(golden_dirs, golden_next, name, capsys):
    """data_loader.py:161-204 (groups, member order, state strings)."""
    from utility.utility_data.data_loader import Data
    cfg = {"dataset": name, "sparsity_test": "1", "top_K": "[20, 40]"}
    d = Data(golden_dirs[name], cfg)
    ptr, flat = golden_next["split_%s_ptr" % name], golden_next["split_%s_flat" % name]
    assert len(d.split_test_dict) == len(ptr) - 1
    for j, users in enumerate(d.split_test_dict):
        assert list(users) == flat[ptr[j]:ptr[j + 1]].tolist()
    assert d.split_state == golden_next["split_%s_state" % name].tolist()
    # in-memory constructor takes the same path
    d2 = Data.from_arrays(d.num_users, d.num_items, d.train_user, d.train_item, d.test_user, d.test_item, cfg)
    assert [list(u) for u in d2.split_test_dict] == [list(u) for u in d.split_test_dict]
