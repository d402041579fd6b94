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


# ------------------------------------------------------------------------------------------------
# oracle restatements of the added models, pinned against the unmodified reference
# ------------------------------------------------------------------------------------------------
NEXT_CFG = {
    "LightCCF": dict(reg_lambda=1e-4, ssl_lambda=5.0, temperature=0.22),
    "LightCSCF": dict(lambda_reg=1e-4, lambda_gamma=1.0, lambda_margin=0.7, temperature=0.2),
    "SCCF": dict(temperature=0.1),
    "DirectAU": dict(gamma=2.0, reg_lambda=1e-4),
    "SGL": dict(reg_lambda=1e-4, ssl_lambda=0.1, temperature=0.2),
}


def _tiny_graph(golden_dirs):
    from oracle import ref_oracle as O
    d = O.load_dataset(golden_dirs["tiny"])
    ip, ix, dt, _ = O.norm_adjacency(d.user_item_net)
    return d, O.csr_to_torch_coo(ip, ix, dt, d.num_nodes)


def _close(a, b, rtol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("kind", ["LightCCF", "LightCSCF", "SCCF", "DirectAU"])
@pytest.mark.parametrize("enc", ["LightGCN", "MF"])
def test_oracle_next_models_match_reference(golden_dirs, golden_tiny, golden_next, kind, enc):
    from oracle import ref_oracle as O
    d, A = _tiny_graph(golden_dirs)
    b = golden_next["batch"]
    r = O.next_model_step(kind, A, golden_tiny["lg_user_w0"], golden_tiny["lg_item_w0"], b[:, 0], b[:, 1], b[:, 2], NEXT_CFG[kind], enc)
    p = "%s_%s" % (kind.lower(), enc.lower())
    np.testing.assert_allclose(r.losses, golden_next[p + "_loss"], rtol=2e-6)
    _close(r.grad_user, golden_next[p + "_gu"], 1e-5)
    _close(r.grad_item, golden_next[p + "_gi"], 1e-5)


def test_oracle_sgl_matches_reference(golden_dirs, golden_tiny, golden_next):
    from oracle import ref_oracle as O
    d, A = _tiny_graph(golden_dirs)
    subs = []
    for j in range(2):
        ip, ix, dt = O.subgraph_adjacency(d.user_item_net, golden_next["sgl_keep%d" % j])
        np.testing.assert_array_equal(ip, golden_next["sgl_sub%d_indptr" % j])
        np.testing.assert_array_equal(ix, golden_next["sgl_sub%d_indices" % j])
        np.testing.assert_array_equal(dt.view(np.uint32), golden_next["sgl_sub%d_data" % j].view(np.uint32))
        subs.append(O.csr_to_torch_coo(ip, ix, dt, d.num_nodes))
    b = golden_next["batch"]
    r = O.next_model_step("SGL", A, golden_tiny["lg_user_w0"], golden_tiny["lg_item_w0"], b[:, 0], b[:, 1], b[:, 2], NEXT_CFG["SGL"], sub_graphs=subs)
    np.testing.assert_allclose(r.losses, golden_next["sgl_loss"], rtol=2e-6)
    _close(r.grad_user, golden_next["sgl_gu"], 1e-5)
    _close(r.grad_item, golden_next["sgl_gi"], 1e-5)


def test_oracle_sparsity_groups_match_reference(golden_dirs, golden_tiny, golden_next):
    from oracle import ref_oracle as O
    d, A = _tiny_graph(golden_dirs)
    groups = O.sparsity_split(d)
    ptr, flat = golden_next["split_tiny_ptr"], golden_next["split_tiny_flat"]
    assert [list(g) for g in groups] == [flat[ptr[j]:ptr[j + 1]].tolist() for j in range(len(ptr) - 1)]
    assert bool(golden_next["sparsity_ok"])
    om = O.OracleModel("LightGCN", A, golden_tiny["lg_user_wT"], golden_tiny["lg_item_wT"])
    fu, fi = om.final_embeddings()
    res = O.evaluate_groups(fu, fi, d, groups, [20, 40])
    for j, r in enumerate(res):
        for k in ("recall", "precision", "ndcg"):
            np.testing.assert_allclose(r[k], golden_next["sparsity_" + k][j], atol=5e-5)
    full, _ = O.evaluate(fu, fi, d, [20, 40], 37)
    for k in ("recall", "precision", "ndcg"):
        np.testing.assert_allclose(full[k], golden_next["test2040_" + k], atol=5e-5)


@pytest.mark.parametrize("mode", ["parallel", "alternating"])
def test_oracle_egcf_matches_reference(golden_dirs, golden_next, mode):
    from oracle import ref_oracle as O
    d, A = _tiny_graph(golden_dirs)
    Rt = O.bipartite_adjacency(d.user_item_net)
    assert str(golden_next["egcf_R_dtype"]) == "float64"
    np.testing.assert_array_equal(Rt.indices().numpy(), golden_next["egcf_R_index"])
    np.testing.assert_array_equal(Rt.values().numpy().view(np.uint32), golden_next["egcf_R_value"].view(np.uint32))
    b = golden_next["batch"]
    cfg = dict(reg_lambda=1e-4, ssl_lambda=0.1, temperature=0.1)
    losses, gi, fu, fi = O.egcf_step(Rt, A, golden_next["egcf_item_w0"], b[:, 0], b[:, 1], b[:, 2], cfg, mode)
    np.testing.assert_allclose(losses, golden_next["egcf_%s_loss" % mode], rtol=2e-6)
    _close(gi, golden_next["egcf_%s_gi" % mode], 1e-5)
    _close(fu, golden_next["egcf_%s_fu" % mode], 1e-6)
    _close(fi, golden_next["egcf_%s_fi" % mode], 1e-6)


def test_functional_known_answers(golden_next):
    from oracle import ref_oracle as O
    a, b = torch.from_numpy(golden_next["fn_a"]), torch.from_numpy(golden_next["fn_b"])
    assert abs(O.align_loss(a, b).item() - float(golden_next["fn_align"])) < 1e-6
    assert abs(O.uniform_loss(a).item() - float(golden_next["fn_uniform"])) < 1e-6


# ------------------------------------------------------------------------------------------------
# the closed forms csrc/pairloss.cu implements (H matrix + diagonal weights + three products +
# normalisation backward), restated in numpy float64 and checked against autograd of the oracle
# ------------------------------------------------------------------------------------------------
def pair_loss_closed_form(kind, X, Y, p0=0.0, p1=0.0):
    n, d = X.shape
    nx = np.maximum(np.linalg.norm(X, axis=1), 1e-12)
    a = X / nx[:, None]
    if Y is not None:
        ny = np.maximum(np.linalg.norm(Y, axis=1), 1e-12)
        b = Y / ny[:, None]
    eye = np.eye(n)
    w = np.zeros(n)
    gb = None
    if kind in (0, 1):
        S, R = a @ b.T, a @ a.T
        t = S + R

        def phi(z):
            f, df = np.exp(z / p0), np.exp(z / p0) / p0
            if kind == 1:
                e2 = np.exp(np.maximum(z - p1, 0) / p0)
                f, df = f + e2, df + (z > p1) * e2 / p0
            return f, df
        f, df = phi(t)
        T = f.sum(1)
        pf, pdf = phi(np.diag(S))
        ratio = pf / T
        q = -1.0 / (n * (ratio + 1e-5))
        H = (-q * ratio / T)[:, None] * df
        w = q * pdf / T
        loss = np.mean(-np.log(ratio + 1e-5))
        ga = H @ b + H @ a + H.T @ a + w[:, None] * b
        gb = H.T @ a + w[:, None] * a
    elif kind == 2:
        S = a @ b.T
        f = np.exp(S / p0) + np.exp(S * S / p0)
        df = np.exp(S / p0) / p0 + 2 * S * np.exp(S * S / p0) / p0
        tot = f.sum()
        loss = np.log(tot / p1)
        H = df / tot
        ga, gb = H @ b, H.T @ a
    elif kind == 3:
        ip = (a * b).sum(1)
        f = np.exp(ip / p0) + np.exp(ip * ip / p0)
        df = np.exp(ip / p0) / p0 + 2 * ip * np.exp(ip * ip / p0) / p0
        loss = -np.mean(np.log(f))
        w = -(df / f) / n
        ga, gb = w[:, None] * b, w[:, None] * a
    elif kind == 4:
        loss = np.mean(((a - b) ** 2).sum(1))
        ga = 2 * (a - b) / n
        gb = -ga
    else:
        R = a @ a.T
        dg = np.diag(R)
        E = np.exp(-2 * np.maximum(dg[:, None] + dg[None, :] - 2 * R, 0)) * (1 - eye)
        rowT = E.sum(1)
        tot = rowT.sum()
        loss = np.log(tot / (n * (n - 1)))
        ga = (-8.0 / tot) * (rowT[:, None] * a - E @ a)
    gX = (ga - a * (a * ga).sum(1, keepdims=True)) / nx[:, None]
    gY = None if gb is None else (gb - b * (b * gb).sum(1, keepdims=True)) / ny[:, None]
    return loss, gX, gY


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4, 5])
def test_pair_loss_closed_forms_equal_autograd(kind):
    from oracle import ref_oracle as O
    rng = np.random.default_rng(kind)
    n, d = 48, 64
    X, Y = rng.normal(size=(n, d)), rng.normal(size=(n, d))
    X[5] = X[9]                                     # duplicate rows (the same user twice in a batch)
    Y += 0.8 * X                                    # similarities spread over both sides of the margin
    user = torch.from_numpy(rng.integers(0, 30, n))
    pos = torch.from_numpy(rng.integers(0, 40, n))
    tx = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    ty = torch.tensor(Y, dtype=torch.float64, requires_grad=True)
    p0, p1 = 0.2, 0.0
    if kind == 0:
        ref = O.lightccf_na_loss(tx, ty, p0)
    elif kind == 1:
        p1 = 0.3
        ref = O.lightcscf_loss(tx, ty, p0, p1)
    elif kind in (2, 3):
        # SCCF gathers from tables: build tables whose gathered rows are X / Y (duplicates share a row)
        fu = torch.tensor(rng.normal(size=(30, d)), dtype=torch.float64, requires_grad=True)
        fi = torch.tensor(rng.normal(size=(40, d)), dtype=torch.float64, requires_grad=True)
        up_neg, down = O.sccf_losses(fu, fi, user, pos, p0)
        X, Y = fu.detach().numpy()[user.numpy()], fi.detach().numpy()[pos.numpy()]
        p1 = float(len(np.unique(user.numpy())) * len(np.unique(pos.numpy())))
        ref = down if kind == 2 else up_neg
        ref.backward()
        loss, gX, gY = pair_loss_closed_form(kind, X, Y, p0, p1)
        assert abs(loss - ref.item()) < 1e-9 * max(1, abs(ref.item()))
        gu = np.zeros((30, d)); np.add.at(gu, user.numpy(), gX)
        gi = np.zeros((40, d)); np.add.at(gi, pos.numpy(), gY)
        np.testing.assert_allclose(gu, fu.grad.numpy(), rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(gi, fi.grad.numpy(), rtol=1e-8, atol=1e-12)
        return
    elif kind == 4:
        ref = O.align_loss(tx, ty)
    else:
        ref = O.uniform_loss(tx)
    ref.backward()
    loss, gX, gY = pair_loss_closed_form(kind, X, None if kind == 5 else Y, p0, p1)
    assert abs(loss - ref.item()) < 1e-9 * max(1, abs(ref.item()))
    np.testing.assert_allclose(gX, tx.grad.numpy(), rtol=1e-7, atol=1e-11)
    if gY is not None:
        np.testing.assert_allclose(gY, ty.grad.numpy(), rtol=1e-7, atol=1e-11)


# ------------------------------------------------------------------------------------------------
# host parser of the dataset text format (idg_parse_ratings)
# ------------------------------------------------------------------------------------------------
def test_parse_ratings_matches_reference_semantics(tmp_path, golden_dirs, golden_quirks):
    from idgrec import _lib, ops
    from oracle import ref_oracle as O
    # the quirks file: duplicate pairs, a user with an empty line, an item id that only occurs in test
    path = os.path.join(golden_dirs["quirks"], "train.txt")
    line_user, line_len, users, items, mu, mi = ops.parse_ratings(path)
    np.testing.assert_array_equal(users, golden_quirks["train_user"])
    np.testing.assert_array_equal(items, golden_quirks["train_item"])
    ref = O.read_ratings(path)
    np.testing.assert_array_equal(line_user, np.asarray(ref[0]))
    assert line_len.sum() == len(items) and (line_len == 0).sum() == 1
    assert mu == users.max() and mi == items.max()
    # tabs / CRLF / several blanks / no trailing newline / blank lines are tolerated, ids may be large
    p = tmp_path / "a.txt"
    p.write_bytes(b"3 1  2\t5\r\n\n7\n0 16777217 4")
    line_user, line_len, users, items, mu, mi = ops.parse_ratings(str(p))
    assert line_user.tolist() == [3, 7, 0] and line_len.tolist() == [3, 0, 2]
    assert users.tolist() == [3, 3, 3, 0, 0] and items.tolist() == [1, 2, 5, 16777217, 4]
    assert (mu, mi) == (3, 16777217)
    # empty file, missing file, junk
    e = tmp_path / "e.txt"
    e.write_bytes(b"")
    assert [len(x) for x in ops.parse_ratings(str(e))[:4]] == [0, 0, 0, 0]
    with pytest.raises(_lib.IdgError):
        ops.parse_ratings(str(tmp_path / "missing.txt"))
    j = tmp_path / "j.txt"
    j.write_bytes(b"1 2 x3\n")
    with pytest.raises(_lib.IdgError):
        ops.parse_ratings(str(j))


@pytest.mark.parametrize("name", ["tiny", "quirks"])
def test_remaining_data_samplers_match_reference(golden_dirs, name):
    """Data.sample_data_to_train_random / get_user_n_neg_items (data_loader.py:86-106,135-149): values and the numpy
    generator state they leave equal the unmodified reference's (tests/golden/samplers.npz)."""
    from utility.utility_data.data_loader import Data
    g = np.load(os.path.join(GOLDEN, "samplers.npz"), allow_pickle=False)
    d = Data(golden_dirs[name], {})
    np.random.seed(77)
    a = d.sample_data_to_train_random()
    b = d.get_user_n_neg_items(list(range(0, d.num_users, 3)), 4)
    st = np.random.get_state()
    np.testing.assert_array_equal(a, g["rand_" + name])
    np.testing.assert_array_equal(np.array(b, dtype=np.int64), g["nneg_" + name])
    np.testing.assert_array_equal(st[1], g["rng_key_" + name])
    assert st[2] == int(g["rng_pos_" + name])
