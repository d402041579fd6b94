"""Generates tests/golden/samplers.npz by running the UNMODIFIED reference Data class (build container only):
sample_data_to_train_random (data_loader.py:86-106) and get_user_n_neg_items (:135-149) on the committed tiny / quirks
datasets from np.random.seed(77), together with the generator state they leave behind.
    python tests/golden/make_golden_samplers.py"""
import os
import shutil
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, "/root/reference")
import utility.utility_data.data_loader as ref  # noqa: E402

assert ref.__file__.startswith("/root/reference")


def main():
    root = tempfile.mkdtemp(prefix="idgrec_golden_samplers_")
    res = {}
    try:
        for n in ("tiny", "quirks"):
            os.makedirs(os.path.join(root, n))
            shutil.copy(os.path.join(OUT, n + "_train.txt"), os.path.join(root, n, "train.txt"))
            shutil.copy(os.path.join(OUT, n + "_test.txt"), os.path.join(root, n, "test.txt"))
            d = ref.Data(os.path.join(root, n), {})
            np.random.seed(77)
            res["rand_" + n] = d.sample_data_to_train_random()
            res["nneg_" + n] = np.array(d.get_user_n_neg_items(list(range(0, d.num_users, 3)), 4), dtype=np.int64)
            st = np.random.get_state()
            res["rng_key_" + n], res["rng_pos_" + n] = st[1].copy(), st[2]
        np.savez_compressed(os.path.join(OUT, "samplers.npz"), **res)
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
