"""Property tests (hypothesis) of the host-side C entry points and the vectorised host logic against literal Python
restatements of the reference: dataset text parser, exact negative-sampler walk (chunked candidate stream), activity split."""
import os

import numpy as np
import scipy.sparse as sp
from hypothesis import given, settings, strategies as st

from oracle import ref_oracle as O


# ---- idg_parse_ratings vs the reference's line loop (data_loader.py:48-70) -------------------------------------------
@st.composite
def rating_files(draw):
    n_lines = draw(st.integers(0, 12))
    lines = []
    for _ in range(n_lines):
        user = draw(st.integers(0, 50))
        items = draw(st.lists(st.integers(0, 300), min_size=0, max_size=6))
        lines.append(" ".join(str(x) for x in [user] + items))
    return "\n".join(lines) + ("\n" if draw(st.booleans()) and lines else "")


@settings(max_examples=60, deadline=None)
@given(rating_files())
def test_parser_equals_reference_line_loop(tmp_path_factory, text):
    from idgrec import ops
    p = tmp_path_factory.mktemp("pr") / "r.txt"
    p.write_text(text)
    line_user, line_len, users, items, mu, mi = ops.parse_ratings(str(p))
    # the reference loop, restated (strip, split on blanks, first token = user)
    ru, ri, rl, rlen, max_u, max_i = [], [], [], [], -1, -1
    for line in text.split("\n"):
        tok = line.split()
        if not tok:
            continue
        arr = [int(t) for t in tok]
        rl.append(arr[0]); rlen.append(len(arr) - 1)
        if len(arr) > 1:
            max_u, max_i = max(max_u, arr[0]), max(max_i, max(arr[1:]))
            ru += [arr[0]] * (len(arr) - 1); ri += arr[1:]
    assert users.tolist() == ru and items.tolist() == ri
    assert line_user.tolist() == rl and line_len.tolist() == rlen
    assert (mu, mi) == (max_u, max_i)


# ---- sampler: Data.sample_negatives (chunked stream + C walk) vs the reference's per-edge scalar draws ---------------
@settings(max_examples=25, deadline=None)
@given(st.integers(2, 12), st.integers(3, 15), st.integers(0, 2 ** 31 - 1), st.floats(0.05, 0.8))
def test_sampler_equals_scalar_draw_loop(U, I, seed, density):
    from utility.utility_data.data_loader import Data
    rng = np.random.default_rng(seed)
    mask = rng.random((U, I)) < density
    mask[:, 0] = False                      # every user keeps at least one non-positive item, or the reference loops forever
    tu, ti = np.nonzero(mask)
    if len(tu) == 0:
        return
    order = rng.permutation(len(tu))        # file order need not be sorted
    tu, ti = tu[order], ti[order]
    d = Data.from_arrays(U, I, tu, ti, tu[:1], ti[:1], {})
    np.random.seed(seed % 1000)
    neg = d.sample_negatives()
    state_a = np.random.get_state()
    # literal restatement of data_loader.py:108-127
    np.random.seed(seed % 1000)
    want = []
    for e in range(len(tu)):
        pos = d.all_positive[tu[e]]
        while True:
            c = np.random.randint(0, I)
            if c in pos:
                continue
            break
        want.append(c)
    state_b = np.random.get_state()
    assert neg.tolist() == want
    assert state_a[2] == state_b[2] and np.array_equal(state_a[1], state_b[1])


# ---- activity split: vectorised Data.create_sparsity_split vs the statement-by-statement oracle ----------------------
@settings(max_examples=40, deadline=None)
@given(st.integers(2, 25), st.integers(4, 30), st.integers(0, 2 ** 31 - 1))
def test_sparsity_split_equals_literal_restatement(U, I, seed):
    from utility.utility_data.data_loader import Data
    rng = np.random.default_rng(seed)
    mask = rng.random((U, I)) < rng.uniform(0.1, 0.7)
    tu, ti = np.nonzero(mask)
    su, si = np.nonzero(~mask & (rng.random((U, I)) < 0.3))
    if len(tu) == 0 or len(su) == 0:
        return
    d = Data.from_arrays(U, I, tu, ti, su, si, {"sparsity_test": "1"})
    od = O.OracleData(path="", num_users=U, num_items=I, num_nodes=U + I, num_train=len(tu), num_test=len(su), train_user=tu, train_item=ti,
                      test_user=su, test_item=si, user_item_net=d.user_item_net, all_positive=d.all_positive, test_dict=d.test_dict)
    assert [list(g) for g in d.split_test_dict] == [list(g) for g in O.sparsity_split(od)]
