"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle and the golden
vectors of the unmodified reference.  Run on the B200 box: ``pytest -m gpu``."""
import numpy as np
import pytest
import torch

from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star: embeddings and losses within 1e-5 relative (fp32)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from idgrec import _lib
    _lib.lib()  # fails loudly if the extension is missing
    return torch.device("cuda:0")


def _tiny(golden_dirs):
    return O.load_dataset(golden_dirs["tiny"])


def _rand_graph(U, I, E, seed, hub=0):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, U, E)
    i = (rng.zipf(1.3, E) - 1) % I
    if hub:  # one user and one item far above the 256-nonzero chunk size
        u = np.concatenate([u, np.zeros(hub, np.int64), rng.integers(0, U, hub)])
        i = np.concatenate([i, rng.permutation(I)[:hub], np.zeros(hub, np.int64)])
    key = np.unique(u.astype(np.int64) * I + i)
    return key // I, key % I


def _assert_close(a, b, rtol=RTOL, scale=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.abs(b).max() if scale is None else scale
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * max(s, 1e-30))


# ---------------------------------------------------------------- a2: CSR build
@pytest.mark.parametrize("name", ["tiny", "quirks"])
@pytest.mark.parametrize("add_self", [False, True])
def test_csr_build_bit_exact_golden(dev, golden_dirs, golden_tiny, golden_quirks, name, add_self):
    from idgrec.graph import build_norm_adjacency
    g = golden_tiny if name == "tiny" else golden_quirks
    d = O.load_dataset(golden_dirs[name])
    csr = build_norm_adjacency(d.train_user, d.train_item, d.num_users, d.num_items, add_self=add_self, device=dev)
    indptr, indices, data = csr.indptr.cpu().numpy(), csr.indices.cpu().numpy(), csr.data.cpu().numpy()
    rows = np.repeat(np.arange(d.num_nodes), np.diff(indptr))
    pre = "As" if add_self else "A"
    np.testing.assert_array_equal(np.stack([rows, indices]), g[pre + "_coo_index"])
    np.testing.assert_array_equal(data.view(np.uint32), g[pre + "_coo_value"].view(np.uint32))
    if not add_self:
        np.testing.assert_array_equal(indptr, g["A_indptr"])


@pytest.mark.parametrize("add_self", [False, True])
def test_csr_build_bit_exact_random(dev, add_self):
    from idgrec.graph import build_norm_adjacency
    import scipy.sparse as sp
    U, I = 3000, 4100
    u, i = _rand_graph(U, I, 90000, 7, hub=700)
    net = sp.csr_matrix((np.ones(len(u)), (u, i)), shape=(U, I))
    ip, ix, dt, _ = O.norm_adjacency(net, add_self=add_self)
    csr = build_norm_adjacency(u, i, U, I, add_self=add_self, device=dev)
    np.testing.assert_array_equal(csr.indptr.cpu().numpy(), ip)
    np.testing.assert_array_equal(csr.indices.cpu().numpy(), ix)
    np.testing.assert_array_equal(csr.data.cpu().numpy().view(np.uint32), dt.view(np.uint32))


# ---------------------------------------------------------------- a6/a7/a11: propagation
def _graph_and_oracle(dev, U, I, u, i, add_self=False):
    from idgrec.graph import build_norm_adjacency, Graph
    import scipy.sparse as sp
    net = sp.csr_matrix((np.ones(len(u)), (u, i)), shape=(U, I))
    ip, ix, dt, _ = O.norm_adjacency(net, add_self=add_self)
    A = O.csr_to_torch_coo(ip, ix, dt, U + I)
    csr = build_norm_adjacency(u, i, U, I, add_self=add_self, device=dev)
    return Graph(csr), A


@pytest.mark.parametrize("d", [32, 64, 128])
def test_spmm_layer_vs_sparse_mm(dev, d):
    U, I = 1500, 2100
    u, i = _rand_graph(U, I, 40000, 3, hub=900)
    G, A = _graph_and_oracle(dev, U, I, u, i)
    X = torch.randn(U + I, d, generator=torch.Generator().manual_seed(1))
    ref = torch.sparse.mm(A, X).numpy()
    Xd = X.to(dev)
    Y = torch.empty_like(Xd)
    G.spmm_layer(Xd, Y=Y)
    _assert_close(Y.cpu().numpy(), ref)
    # determinism: two launches give identical bits
    Y2 = torch.empty_like(Xd)
    G.spmm_layer(Xd, Y=Y2)
    assert torch.equal(Y, Y2)
    # epilogues: addend, running layer sum with division
    add = torch.randn_like(Xd)
    acc = torch.empty_like(Xd)
    G.spmm_layer(Xd, Y=None, addend=add, acc_in=Xd, acc_out=acc, acc_div=4.0)
    _assert_close(acc.cpu().numpy(), ((X.numpy() + ref + add.cpu().numpy()) / 4.0))


@pytest.mark.parametrize("inc0,use_noise,cl", [(True, False, 0), (False, False, 0), (False, True, 0), (False, True, 1), (False, True, 2), (False, True, 3)])
def test_propagate_fwd_bwd(dev, inc0, use_noise, cl):
    from idgrec import ops
    U, I, K, d = 900, 1300, 3, 64
    u, i = _rand_graph(U, I, 25000, 11, hub=500)
    G, A = _graph_and_oracle(dev, U, I, u, i)
    gen = torch.Generator().manual_seed(5)
    X0 = (torch.rand(U + I, d, generator=gen) - 0.5) * 0.2
    noises = [torch.rand(U + I, d, generator=gen) for _ in range(K)] if use_noise else None
    eps = 0.1
    Xr = X0.clone().requires_grad_(True)
    out = O.propagate(A, Xr, K, inc0, noises, eps, cl)
    Xg = X0.to(dev).requires_grad_(True)
    nz = torch.stack(noises).to(dev).contiguous() if use_noise else None
    res = ops.propagate(Xg, G, K, inc0, noise=nz, eps=eps, cl_layer=cl)
    wF = torch.randn(U + I, d, generator=gen)
    wC = torch.randn(U + I, d, generator=gen)
    if cl > 0:
        _assert_close(res[0].detach().cpu().numpy(), out[0].detach().numpy())
        _assert_close(res[1].detach().cpu().numpy(), out[1].detach().numpy())
        ((out[0] * wF).sum() + (out[1] * wC).sum()).backward()
        ((res[0] * wF.to(dev)).sum() + (res[1] * wC.to(dev)).sum()).backward()
    else:
        _assert_close(res.detach().cpu().numpy(), out.detach().numpy())
        (out * wF).sum().backward()
        (res * wF.to(dev)).sum().backward()
    _assert_close(Xg.grad.cpu().numpy(), Xr.grad.numpy())


def test_row_restricted_and_sparse_input_layers(dev):
    """The two identical-result shortcuts of the fused trainer: last forward layer only on the batch
    rows, first backward product only over the batch columns (heavy rows included)."""
    from idgrec.graph import BatchRows
    U, I, K, d, B = 900, 1300, 3, 64, 300
    u, i = _rand_graph(U, I, 25000, 13, hub=600)
    G, A = _graph_and_oracle(dev, U, I, u, i)
    gen = torch.Generator().manual_seed(9)
    X0 = ((torch.rand(U + I, d, generator=gen) - 0.5) * 0.2).to(dev)
    users = torch.randint(0, U, (B,), generator=gen)
    users[:5] = 0                                   # hub user (heavy row)
    pos = torch.randint(0, 30, (B,), generator=gen)
    pos[:5] = 0                                     # hub item
    neg = torch.randint(0, I, (B,), generator=gen)
    rows = BatchRows(U + I, B, dev)
    ud, pd_, nd = users.to(dev), pos.to(dev), neg.to(dev)
    rows.build(ud.data_ptr(), pd_.data_ptr(), nd.data_ptr(), B, U)
    want = np.unique(np.concatenate([users.numpy(), U + pos.numpy(), U + neg.numpy()]))
    n = int(rows.count.item())
    assert n == len(want)
    np.testing.assert_array_equal(np.sort(rows.rowlist[:n].cpu().numpy()), want)
    bits = rows.bitmap.cpu().numpy().view(np.uint32)
    got = np.flatnonzero(np.unpackbits(bits.view(np.uint8), bitorder="little"))
    np.testing.assert_array_equal(got, want)
    # forward: restricted rows equal the full propagation bit for bit
    full = G.propagate_fwd(X0, K, True)
    part = G.propagate_fwd(X0, K, True, rows=rows)
    w = torch.from_numpy(want).to(dev)
    assert torch.equal(full[w], part[w])
    # backward: G non-zero only on the batch rows
    Gd = torch.zeros(U + I, d, device=dev)
    Gd[w] = torch.randn(len(want), d, generator=gen).to(dev)
    dense = G.propagate_bwd(Gd, K, True)
    sparse = G.propagate_bwd(Gd, K, True, rows=rows)
    _assert_close(sparse.cpu().numpy(), dense.cpu().numpy(), rtol=1e-6)
    ref = torch.from_numpy(Gd.cpu().numpy())
    h = ref
    for _ in range(K):
        h = ref + torch.sparse.mm(A, h)
    _assert_close(sparse.cpu().numpy(), (h / (K + 1)).numpy())
    # neighbourhood ("closure") restriction: layer K-1 only on batch rows + neighbours, second backward product over them
    rows.enable_closure(G)
    rows.build_closure(G)
    import scipy.sparse as sp
    Acsr = G.csr.to_scipy()
    nb = np.zeros(U + I, bool)
    nb[want] = True
    nb[np.unique(Acsr[want].indices)] = True   # A_hat is symmetric: neighbours of the batch rows
    cbits = rows.closure.cpu().numpy().view(np.uint32)
    got_c = np.flatnonzero(np.unpackbits(cbits.view(np.uint8), bitorder="little"))
    np.testing.assert_array_equal(got_c, np.flatnonzero(nb))
    part2 = G.propagate_fwd(X0, K, True, rows=rows)
    assert torch.equal(full[w], part2[w])
    sparse2 = G.propagate_bwd(Gd, K, True, rows=rows)
    _assert_close(sparse2.cpu().numpy(), dense.cpu().numpy(), rtol=1e-6)
    from idgrec import _lib
    _lib.check(_lib.lib().idg_graph_set_closure(G._h, None))
    rows.clear()
    assert int(rows.bitmap.abs().sum().item()) == 0 and int(rows.closure.abs().sum().item()) == 0


# ---------------------------------------------------------------- a9: BPR + reg
@pytest.mark.parametrize("B,reg_mask", [(256, 7), (1000, 7), (333, 6)])
def test_bpr_reg_loss_and_grads(dev, B, reg_mask):
    from idgrec import ops
    U, I, d = 300, 500, 64
    gen = torch.Generator().manual_seed(B)
    F = torch.randn(U + I, d, generator=gen) * 0.3
    E0 = torch.randn(U + I, d, generator=gen) * 0.1
    users = torch.randint(0, U, (B,), generator=gen)
    pos = torch.randint(0, 40, (B,), generator=gen)  # many duplicates -> exercises the ordered scatter
    neg = torch.randint(0, I, (B,), generator=gen)
    Fr, Er = F.clone().requires_grad_(True), E0.clone().requires_grad_(True)
    fu, fi = Fr[:U], Fr[U:]
    bpr = O.bpr_loss(fu[users], fi[pos], fi[neg])
    embs = [Er[users], Er[U + pos], Er[U + neg]]
    reg = 1e-4 * O.reg_loss(*[e for k, e in enumerate(embs) if (reg_mask >> k) & 1])
    (bpr * 1.0 + reg * 1.0).backward()
    Fg, Eg = F.to(dev).requires_grad_(True), E0.to(dev).requires_grad_(True)
    loss = ops.bpr_reg_loss(Fg, Eg, users.to(dev), pos.to(dev), neg.to(dev), U, 1e-4, reg_mask)
    np.testing.assert_allclose(loss.detach().cpu().numpy(), [bpr.item(), reg.item()], rtol=RTOL)
    loss.sum().backward()
    _assert_close(Fg.grad.cpu().numpy(), Fr.grad.numpy())
    _assert_close(Eg.grad.cpu().numpy(), Er.grad.numpy())


def test_functional_known_answers_gpu(dev, golden_tiny):
    """losses.py known answers produced by the reference itself (fn_* in tiny.npz)."""
    from idgrec import ops
    g = golden_tiny
    a, b, c = (torch.from_numpy(g[k]).to(dev) for k in ("fn_a", "fn_b", "fn_c"))
    n = a.shape[0]
    F = torch.cat([a, b, c])  # "users" = rows of a, "items" = rows of b then c
    idx = torch.arange(n, device=dev)
    loss = ops.bpr_reg_loss(F, F, idx, idx, idx + n, n, 1.0, 7)
    np.testing.assert_allclose(loss[0].item(), float(g["fn_bpr"]), rtol=RTOL)
    np.testing.assert_allclose(loss[1].item(), float(g["fn_reg"]), rtol=RTOL)
    nce = ops.infonce_rows(torch.cat([a, a]), torch.cat([b, b]), idx, 0.2)
    np.testing.assert_allclose(nce.item(), float(g["fn_nce"]), rtol=RTOL)


# ---------------------------------------------------------------- a10: InfoNCE
@pytest.mark.parametrize("n,tau", [(37, 0.2), (200, 0.15), (1111, 0.2)])
def test_infonce_loss_and_grads(dev, n, tau):
    from idgrec import ops
    N, d = n + 50, 64
    gen = torch.Generator().manual_seed(n)
    V1 = torch.randn(N, d, generator=gen)
    V2 = V1 + 0.3 * torch.randn(N, d, generator=gen)
    idx = torch.sort(torch.randperm(N, generator=gen)[:n]).values
    r1, r2 = V1.clone().requires_grad_(True), V2.clone().requires_grad_(True)
    ref = O.infonce_loss(r1[idx], r2[idx], tau)
    ref.backward()
    g1, g2 = V1.to(dev).requires_grad_(True), V2.to(dev).requires_grad_(True)
    out = ops.infonce_rows(g1, g2, idx.to(dev), tau)
    np.testing.assert_allclose(out.item(), ref.item(), rtol=RTOL)
    (out * 1.0).backward()
    _assert_close(g1.grad.cpu().numpy(), r1.grad.numpy(), rtol=5e-5)
    _assert_close(g2.grad.cpu().numpy(), r2.grad.numpy(), rtol=5e-5)


def test_unique_rows_matches_torch_unique(dev):
    from idgrec import _lib
    l = _lib.lib()
    gen = torch.Generator().manual_seed(4)
    for n, hi in ((1, 5), (257, 40), (2048, 900), (4096, 100000)):
        ids = torch.randint(0, hi, (n,), generator=gen).to(dev)
        out = torch.zeros(n, dtype=torch.int64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(l.idg_unique_rows(ids.data_ptr(), n, 7, out.data_ptr(), cnt.data_ptr(), torch.cuda.current_stream().cuda_stream))
        want = torch.unique(ids) + 7
        assert int(cnt.item()) == want.numel()
        assert torch.equal(out[:want.numel()], want)


# ---------------------------------------------------------------- a12: Adam
def test_adam_matches_torch(dev):
    from idgrec import ops
    gen = torch.Generator().manual_seed(0)
    p0 = torch.randn(1000, 64, generator=gen)
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pr], lr=1e-3)
    p, m, v = p0.to(dev), torch.zeros(1000, 64, device=dev), torch.zeros(1000, 64, device=dev)
    for step in range(1, 4):
        g = torch.randn(1000, 64, generator=gen) * 0.01
        pr.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.to(dev), m, v, 1e-3, step)
        np.testing.assert_allclose(p.cpu().numpy(), pr.detach().numpy(), rtol=1e-6, atol=1e-9)


# ---------------------------------------------------------------- a13/a14/a15: evaluation
def _mask_csr(net, dev):
    return (torch.from_numpy(net.indptr.astype(np.int32)).to(dev), torch.from_numpy(net.indices.astype(np.int32)).to(dev))


@pytest.mark.parametrize("scale", [0.1, 0.4, 0.6])
def test_eval_topk_exact_rank(dev, scale):
    """T0: ids bit-exact against the fp64 exact-rank oracle (score desc, id asc)."""
    from idgrec import ops
    import scipy.sparse as sp
    U, I, d, K = 700, 3000, 64, 20
    u, i = _rand_graph(U, I, 30000, 21, hub=400)
    net = sp.csr_matrix((np.ones(len(u)), (u, i)), shape=(U, I))
    net.sort_indices()
    gen = torch.Generator().manual_seed(3)
    Fu = (torch.randn(U, d, generator=gen) * scale).numpy()
    Fi = (torch.randn(I, d, generator=gen) * scale).numpy()
    Fi[100:110] = Fi[90:100]  # exact duplicates -> exact score ties, broken by id
    users = np.arange(U, dtype=np.int64)[::-1].copy()
    ref_ids, ref_sc = O.topk_exact(Fu, Fi, users, net.indptr, net.indices, K)
    mp, mi = _mask_csr(net, dev)
    ids, sc = ops.eval_topk(torch.from_numpy(Fu).to(dev), torch.from_numpy(Fi).to(dev), torch.from_numpy(users).to(dev), mp, mi, K,
                            want_scores=True)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_ids)
    np.testing.assert_allclose(sc.cpu().numpy(), ref_sc, rtol=1e-6)


def test_eval_topk_degenerate_ties_and_short_rows(dev):
    """All-equal scores (zero embeddings) and users with fewer than K unmasked items go
    through the exhaustive pass and still follow (score desc, id asc)."""
    from idgrec import ops
    import scipy.sparse as sp
    U, I, d, K = 70, 300, 64, 20
    rows, cols = [], []
    for uu in range(U):
        m = I - 5 if uu == 3 else (uu % 40)
        rows += [uu] * m
        cols += list(np.random.default_rng(uu).permutation(I)[:m])
    net = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(U, I))
    net.sort_indices()
    Fu = np.zeros((U, d), np.float32)
    Fi = np.zeros((I, d), np.float32)
    Fu[10:] = np.random.default_rng(0).normal(size=(U - 10, d)).astype(np.float32)
    Fi[:] = np.random.default_rng(1).normal(size=(I, d)).astype(np.float32)
    Fi[50:250] = Fi[50]  # 200 identical items
    users = np.arange(U, dtype=np.int64)
    ref_ids, _ = O.topk_exact(Fu, Fi, users, net.indptr, net.indices, K)
    mp, mi = _mask_csr(net, dev)
    ids = ops.eval_topk(torch.from_numpy(Fu).to(dev), torch.from_numpy(Fi).to(dev), torch.from_numpy(users).to(dev), mp, mi, K)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_ids)


def test_eval_metrics_vs_reference(dev, golden_dirs, golden_tiny):
    """T2: device top-K + device metrics reproduce the reference's own Test() numbers."""
    from idgrec import ops
    g, d = golden_tiny, _tiny(golden_dirs)
    users = np.array(list(d.test_dict.keys()), dtype=np.int64)
    tptr = np.zeros(d.num_users + 1, np.int32)
    tl = []
    for uu in range(d.num_users):
        t = sorted(d.test_dict.get(uu, []))
        tl += t
        tptr[uu + 1] = tptr[uu] + len(t)
    mp, mi = _mask_csr(d.user_item_net, dev)
    for tag in ("lgT",):
        Fu, Fi = torch.from_numpy(g[tag + "_fu"]).to(dev), torch.from_numpy(g[tag + "_fi"]).to(dev)
        ids = ops.eval_topk(Fu, Fi, torch.from_numpy(users).to(dev), mp, mi, 20)
        sums = ops.eval_metric_sums(ids, torch.from_numpy(users).to(dev), torch.from_numpy(tptr).to(dev),
                                    torch.tensor(tl, dtype=torch.int32, device=dev), [10, 20]).cpu().numpy() / len(users)
        np.testing.assert_allclose(sums[:, 0], g[tag + "_test_recall"], atol=5e-5)
        np.testing.assert_allclose(sums[:, 1], g[tag + "_test_precision"], atol=5e-5)
        np.testing.assert_allclose(sums[:, 2], g[tag + "_test_ndcg"], atol=5e-5)
        # and the oracle's metric code on the same ids, to float64 rounding
        r = O.hit_matrix(ids.cpu().numpy(), [d.test_dict[int(uu)] for uu in users])
        tlen = np.array([len(d.test_dict[int(uu)]) for uu in users], dtype=np.float64)
        for j, k in enumerate((10, 20)):
            np.testing.assert_allclose(sums[j] * len(users), O.metric_sums(r, tlen, k), rtol=1e-12)


# ---------------------------------------------------------------- a4: sampler replay (host entry point)
def test_neg_sampler_replay_matches_reference_stream(dev, golden_dirs, golden_tiny):
    from idgrec import ops
    g, d = golden_tiny, _tiny(golden_dirs)
    np.random.seed(2024)
    st = np.random.get_state()
    cand = np.random.randint(0, d.num_items, size=len(d.train_user) * 2)
    neg, used = ops.neg_sample_replay(d.train_user, d.user_item_net.indptr, d.user_item_net.indices, cand)
    np.testing.assert_array_equal(neg, g["sample_ep0"][:, 2])
    np.random.set_state(st)
    np.random.randint(0, d.num_items, size=used)
    np.testing.assert_array_equal(O.shuffle_indices(len(neg)), g["perm_ep0"])
