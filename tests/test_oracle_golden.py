"""Pins oracle/ref_oracle.py against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ref_oracle as O


def _data(golden_dirs, name):
    return O.load_dataset(golden_dirs[name])


@pytest.mark.parametrize("name", ["tiny", "quirks"])
def test_loader_matches_reference(golden_dirs, golden_tiny, golden_quirks, name):
    g = golden_tiny if name == "tiny" else golden_quirks
    d = _data(golden_dirs, name)
    assert (d.num_users, d.num_items, d.num_train, d.num_test) == tuple(int(g[k]) for k in ("num_users", "num_items", "num_train", "num_test"))
    np.testing.assert_array_equal(d.train_user, g["train_user"])
    np.testing.assert_array_equal(d.train_item, g["train_item"])
    ptr = g["allpos_ptr"]
    for u in range(d.num_users):
        np.testing.assert_array_equal(d.all_positive[u], g["allpos_flat"][ptr[u]:ptr[u + 1]])
    assert list(d.test_dict.keys()) == g["test_users"].tolist()
    tp = g["test_ptr"]
    for j, u in enumerate(g["test_users"].tolist()):
        assert d.test_dict[u] == g["test_flat"][tp[j]:tp[j + 1]].tolist()
    np.testing.assert_array_equal(d.user_item_net.indptr, g["net_indptr"])
    np.testing.assert_array_equal(d.user_item_net.indices, g["net_indices"])
    np.testing.assert_array_equal(d.user_item_net.data, g["net_data"])


@pytest.mark.parametrize("name", ["tiny", "quirks"])
def test_adjacency_bit_exact(golden_dirs, golden_tiny, golden_quirks, name):
    g = golden_tiny if name == "tiny" else golden_quirks
    d = _data(golden_dirs, name)
    indptr, indices, data, _ = O.norm_adjacency(d.user_item_net, add_self=False)
    assert str(g["A_dtype"]) == "float32"
    np.testing.assert_array_equal(indptr, g["A_indptr"])
    np.testing.assert_array_equal(indices, g["A_indices"])
    assert data.dtype == np.float32
    np.testing.assert_array_equal(data.view(np.uint32), g["A_data"].view(np.uint32))
    # coalesced COO the reference models actually hold
    coo = O.csr_to_torch_coo(indptr, indices, data, d.num_nodes)
    np.testing.assert_array_equal(coo.indices().numpy(), g["A_coo_index"])
    np.testing.assert_array_equal(coo.values().numpy().view(np.uint32), g["A_coo_value"].view(np.uint32))
    # with-self variant: float64 arithmetic, fp32 only after tools.py:101
    indptr, indices, data, _ = O.norm_adjacency(d.user_item_net, add_self=True)
    assert str(g["As_dtype"]) == "float64"
    coo = O.csr_to_torch_coo(indptr, indices, data, d.num_nodes)
    np.testing.assert_array_equal(coo.indices().numpy(), g["As_coo_index"])
    np.testing.assert_array_equal(coo.values().numpy().view(np.uint32), g["As_coo_value"].view(np.uint32))


@pytest.mark.parametrize("name", ["tiny", "quirks"])
@pytest.mark.parametrize("bulk", [False, True])
def test_sampler_and_shuffle_stream(golden_dirs, golden_tiny, golden_quirks, name, bulk):
    g = golden_tiny if name == "tiny" else golden_quirks
    d = _data(golden_dirs, name)
    np.random.seed(2024)
    for ep in range(2):
        s = O.sample_negatives_bulk(d) if bulk else O.sample_negatives(d)
        np.testing.assert_array_equal(s, g["sample_ep%d" % ep])
        np.testing.assert_array_equal(O.shuffle_indices(len(s)), g["perm_ep%d" % ep])
    st = np.random.get_state()
    assert st[2] == int(g["rng_after_pos"])
    np.testing.assert_array_equal(st[1], g["rng_after_key"])


def _model(g, d, kind, prefix, **kw):
    indptr, indices, data, _ = O.norm_adjacency(d.user_item_net)
    A = O.csr_to_torch_coo(indptr, indices, data, d.num_nodes)
    return O.OracleModel(kind, A, g[prefix + "_user_w0"], g[prefix + "_item_w0"], **kw)


def test_lightgcn_two_steps(golden_dirs, golden_tiny):
    g, d = golden_tiny, _data(golden_dirs, "tiny")
    m = _model(g, d, "LightGCN", "lg")
    fu, fi = m.final_embeddings()
    np.testing.assert_allclose(fu, g["lg_fu0"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(fi, g["lg_fi0"], rtol=1e-6, atol=1e-8)
    s0 = g["sample_ep0"][g["perm_ep0"]]
    np.testing.assert_array_equal(s0[:256], g["batch"])
    for st in range(2):
        b = s0[st * 256:(st + 1) * 256]
        r = m.step(b[:, 0], b[:, 1], b[:, 2])
        np.testing.assert_allclose(r.losses, g["lg_loss_s%d" % st], rtol=1e-6)
        np.testing.assert_allclose(r.grad_user, g["lg_gu_s%d" % st], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(r.grad_item, g["lg_gi_s%d" % st], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(m.user_w.detach().numpy(), g["lg_user_w_s%d" % st], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(m.item_w.detach().numpy(), g["lg_item_w_s%d" % st], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("kind,prefix,kw", [
    ("SimGCL", "simgcl", dict(ssl_lambda=0.5, temperature=0.2, eps=0.05)),
    ("XSimGCL", "xsimgcl", dict(ssl_lambda=0.2, temperature=0.15, eps=0.2, cl_layer=1)),
    ("XSimGCL", "xsimgcl2", dict(ssl_lambda=0.2, temperature=0.15, eps=0.2, cl_layer=2)),
])
def test_contrastive_models(golden_dirs, golden_tiny, kind, prefix, kw):
    g, d = golden_tiny, _data(golden_dirs, "tiny")
    m = _model(g, d, kind, prefix, **kw)
    noise = [torch.from_numpy(n) for n in g[prefix + "_noise"]]
    noises = [noise[:3], noise[3:6]] if kind == "SimGCL" else noise[:3]
    b = g["batch"]
    r = m.step(b[:, 0], b[:, 1], b[:, 2], noises=noises, apply_adam=False)
    np.testing.assert_allclose(r.losses, g[prefix + "_loss"], rtol=2e-6)
    np.testing.assert_allclose(r.grad_user, g[prefix + "_gu"], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(r.grad_item, g[prefix + "_gi"], rtol=1e-4, atol=1e-8)
    fu, fi = m.final_embeddings()
    np.testing.assert_allclose(fu, g[prefix + "_fu"], rtol=1e-6, atol=1e-8)


def test_ngcf_forward_backward(golden_dirs, golden_tiny):
    g, d = golden_tiny, _data(golden_dirs, "tiny")
    indptr, indices, data, _ = O.norm_adjacency(d.user_item_net, add_self=True)
    A = O.csr_to_torch_coo(indptr, indices, data, d.num_nodes)
    uw = torch.nn.Parameter(torch.from_numpy(g["ngcf_user_w0"].copy()))
    iw = torch.nn.Parameter(torch.from_numpy(g["ngcf_item_w0"].copy()))
    W = {k: [torch.nn.Parameter(torch.from_numpy(g["ngcf_%s_%d" % (k, l)].copy())) for l in range(3)]
         for k in ("W_gcn", "b_gcn", "W_bi", "b_bi")}
    masks = [torch.from_numpy(m) for m in g["ngcf_masks"]]
    F = O.ngcf_aggregate(A, torch.cat([uw, iw]), W["W_gcn"], W["b_gcn"], W["W_bi"], W["b_bi"], masks, [0.1] * 3)
    fu, fi = torch.split(F, [d.num_users, d.num_items])
    b = torch.from_numpy(g["batch"].copy()).long()
    bpr = O.bpr_loss(fu[b[:, 0]], fi[b[:, 1]], fi[b[:, 2]])
    reg = 1e-4 * O.reg_loss(iw[b[:, 1]], iw[b[:, 2]])  # NGCF.py:120-125: item ego rows only
    np.testing.assert_allclose([bpr.item(), reg.item()], g["ngcf_loss"], rtol=2e-6)
    (bpr + reg).backward()
    np.testing.assert_allclose(uw.grad.numpy(), g["ngcf_gu"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(iw.grad.numpy(), g["ngcf_gi"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(W["W_bi"][1].grad.numpy(), g["ngcf_g_W_bi_1"], rtol=1e-4, atol=1e-9)


def test_functional_known_answers(golden_tiny):
    g = golden_tiny
    a, b, c = (torch.from_numpy(g[k]) for k in ("fn_a", "fn_b", "fn_c"))
    np.testing.assert_allclose(O.bpr_loss(a, b, c).item(), float(g["fn_bpr"]), rtol=1e-6)
    np.testing.assert_allclose(O.reg_loss(a, b, c).item(), float(g["fn_reg"]), rtol=1e-6)
    np.testing.assert_allclose(O.infonce_loss(a, b, 0.2).item(), float(g["fn_nce"]), rtol=1e-6)
    truth = [[1, 2, 3], [7], list(range(4, 7)) + list(range(8, 27))]
    pred = np.array([[1, 9, 2, 8, 3], [0, 1, 2, 3, 4], [4, 5, 6, 8, 9]])
    r = O.hit_matrix(pred, truth)
    np.testing.assert_array_equal(r, g["mt_r"])
    tl = np.array([len(t) for t in truth], dtype=np.float64)
    for j, k in enumerate((3, 5)):
        np.testing.assert_allclose(O.metric_sums(r, tl, k), g["mt_vals"][j], rtol=1e-12)


@pytest.mark.parametrize("tag,wu,wi", [("lg", "lg_user_w_s1", "lg_item_w_s1"), ("lgT", "lg_user_wT", "lg_item_wT")])
def test_eval_metrics_vs_reference_Test(golden_dirs, golden_tiny, tag, wu, wi):
    """T2: recall/ndcg of the oracle's exact-rank evaluation equal the reference's
    own Test() to 4 decimals; T1: the reference-faithful ranking reproduces them
    to float64 rounding."""
    g, d = golden_tiny, _data(golden_dirs, "tiny")
    indptr, indices, data, _ = O.norm_adjacency(d.user_item_net)
    A = O.csr_to_torch_coo(indptr, indices, data, d.num_nodes)
    m = O.OracleModel("LightGCN", A, g[wu], g[wi])
    fu, fi = m.final_embeddings()
    for mode, tol in (("reference", 1e-12), ("exact", 5e-5)):
        res, _ = O.evaluate(fu, fi, d, [10, 20], 50, mode=mode)
        for k in ("recall", "precision", "ndcg"):
            np.testing.assert_allclose(res[k], g[tag + "_test_" + k], rtol=0, atol=tol, err_msg=mode + k)


def test_rating_matrix_T1(golden_dirs, golden_tiny):
    g, d = golden_tiny, _data(golden_dirs, "tiny")
    users = g["test_users"][:50]
    _, rating = O.topk_reference_faithful(g["lgT_fu"], g["lgT_fi"], users,
                                          np.zeros(d.num_users + 1, np.int64), np.zeros(0, np.int64), 20)
    np.testing.assert_allclose(rating, g["lgT_rating50"], rtol=1e-6)
