"""GPU parity of the SURVEY section-8(f) rows: sparsity_test evaluator and wider top-K lists on the ranking
kernel, the batch x batch loss kernels (csrc/pairloss.cu) with their deterministic gather / scatter, and the
models built on them (SGL, LightCCF, LightCSCF, SCCF, DirectAU) against outputs of the unmodified reference
(tests/golden/next.npz) and the CPU oracle."""
import importlib
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REPO
from oracle import ref_oracle as O
from test_next_cpu import NEXT_CFG, pair_loss_closed_form

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def golden_next():
    return np.load(os.path.join(GOLDEN, "next.npz"), allow_pickle=False)


def _cfg(name, **over):
    import utility.utility_function.tools as tools
    c = tools.read_configuration(os.path.join(REPO, "id-grec_b200", "configure", name + ".txt"), name)
    c.update(dataset="tiny", **{k: str(v) for k, v in over.items()})
    return c


def _data(golden_dirs, cfg, name="tiny"):
    from utility.utility_data.data_loader import Data
    return Data(golden_dirs[name], cfg)


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * max(np.abs(b).max(), 1e-30))


def _load_weights(model, uw, iw):
    with torch.no_grad():
        model.user_embedding.weight.copy_(torch.from_numpy(uw))
        model.item_embedding.weight.copy_(torch.from_numpy(iw))


# ------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,N", [(1, 64, 5), (37, 64, 20), (256, 64, 90), (1000, 32, 300)])
def test_gather_scatter_rows_with_duplicates(dev, n, d, N):
    """out = T[idx]; backward sums duplicates (index_put accumulate) -- against numpy, and bit-reproducible."""
    from idgrec import ops
    rng = np.random.default_rng(n)
    T = rng.normal(size=(N, d)).astype(np.float32)
    idx = rng.integers(0, N, n)
    G = rng.normal(size=(n, d)).astype(np.float32)
    t = torch.from_numpy(T).to(dev).requires_grad_(True)
    out = ops.gather_rows(t, torch.from_numpy(idx).to(dev))
    np.testing.assert_array_equal(out.detach().cpu().numpy(), T[idx])
    out.backward(torch.from_numpy(G).to(dev))
    want = np.zeros((N, d), np.float64)
    np.add.at(want, idx, G.astype(np.float64))
    got = t.grad.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
    t2 = torch.from_numpy(T).to(dev).requires_grad_(True)
    ops.gather_rows(t2, torch.from_numpy(idx).to(dev)).backward(torch.from_numpy(G).to(dev))
    assert torch.equal(t.grad, t2.grad)


@pytest.mark.parametrize("kind", ["lightccf", "lightcscf", "sccf_down", "sccf_up", "align", "uniform"])
@pytest.mark.parametrize("n,d", [(2, 64), (48, 64), (257, 64), (1000, 32), (1000, 64)])
def test_pair_loss_kernels_vs_closed_form(dev, kind, n, d):
    """Loss and both gradients of every kind against the float64 closed form (itself checked against autograd of
    the oracle in tests/test_next_cpu.py); sizes off the 64-tile grid, duplicate rows, both margin branches."""
    from idgrec import ops
    k = ops.PAIR_KINDS[kind]
    rng = np.random.default_rng(1000 * k + n)
    X = rng.normal(size=(n, d)).astype(np.float32)
    Y = (rng.normal(size=(n, d)) + 0.8 * X).astype(np.float32)
    if n > 10:
        X[5] = X[9]
    p0, p1 = 0.2, 0.0
    if kind == "lightcscf":
        p1 = 0.3
    if kind == "sccf_down":
        p1 = float(max(n * n // 2, 1))
    tx = torch.from_numpy(X).to(dev).requires_grad_(True)
    ty = torch.from_numpy(Y).to(dev).requires_grad_(True)
    loss = ops.pair_loss(kind, tx, None if kind == "uniform" else ty, p0, p1)
    loss.backward()
    ref, gX, gY = pair_loss_closed_form(k, X.astype(np.float64), None if kind == "uniform" else Y.astype(np.float64), p0, p1)
    assert abs(loss.item() - ref) <= 2e-5 * max(1.0, abs(ref)), (loss.item(), ref)
    # the gradient is a difference of O(1/(n tau)) terms; for a 2-row batch they cancel to ~1e-4, so the fp32 rounding
    # of the terms (1e-7 absolute) is the floor of the comparison there
    floor = 1e-7 if n <= 2 else 0.0
    np.testing.assert_allclose(tx.grad.cpu().numpy(), gX, rtol=2e-5, atol=max(floor, 2e-5 * np.abs(gX).max()))
    if gY is not None:
        np.testing.assert_allclose(ty.grad.cpu().numpy(), gY, rtol=2e-5, atol=max(floor, 2e-5 * np.abs(gY).max()))
    # forward-only call (no gradient buffers) gives the same loss, and the result is bit-reproducible
    with torch.no_grad():
        l2 = ops.pair_loss(kind, tx.detach(), None if kind == "uniform" else ty.detach(), p0, p1)
    assert l2.item() == loss.item()


def test_pair_loss_full_batch_size(dev):
    """B = 4096 (configure/LightCCF.txt, LightCSCF.txt): loss against the float64 closed form, gradients through
    a size-independent property: every loss is invariant to the scale of a row, so <x_i, dL/dx_i> = 0."""
    from idgrec import ops
    rng = np.random.default_rng(7)
    n, d = 4096, 64
    X = rng.normal(size=(n, d)).astype(np.float32)
    Y = (rng.normal(size=(n, d)) + 0.5 * X).astype(np.float32)
    for kind, p0, p1 in (("lightccf", 0.22, 0.0), ("lightcscf", 0.2, 0.7), ("uniform", 0.0, 0.0)):
        tx = torch.from_numpy(X).to(dev).requires_grad_(True)
        ty = torch.from_numpy(Y).to(dev).requires_grad_(True)
        loss = ops.pair_loss(kind, tx, None if kind == "uniform" else ty, p0, p1)
        loss.backward()
        ref, gX, _ = pair_loss_closed_form(ops.PAIR_KINDS[kind], X.astype(np.float64), None if kind == "uniform" else Y.astype(np.float64), p0, p1)
        assert abs(loss.item() - ref) <= 2e-5 * max(1.0, abs(ref))
        g = tx.grad
        radial = (g * tx.detach()).sum(1).abs().max().item()
        assert radial <= 1e-4 * float(g.abs().max()) * float(tx.detach().norm(dim=1).max())
        _close(g.cpu().numpy(), gX, rtol=5e-5)


def test_pair_loss_rejects_bad_arguments(dev):
    from idgrec import _lib, ops
    x = torch.zeros(4, 64, device=dev)
    with pytest.raises(_lib.IdgError):
        ops.pair_loss("lightccf", x, x, 0.0)          # temperature must be > 0
    with pytest.raises(_lib.IdgError):
        ops.pair_loss("sccf_down", x, x, 0.1, 0.0)    # needs the unique-count product
    with pytest.raises(_lib.IdgError):
        ops.pair_loss("uniform", x[:1])               # pdist of one row is empty


# ------------------------------------------------------------------------------------------------
# models vs the unmodified reference
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["LightCCF", "LightCSCF", "SCCF", "DirectAU"])
@pytest.mark.parametrize("enc", ["LightGCN", "MF"])
def test_loss_only_models_vs_reference(dev, golden_dirs, golden_tiny, golden_next, kind, enc):
    cfg = _cfg(kind, batch_size=256, encoder=enc, **NEXT_CFG[kind])
    d = _data(golden_dirs, cfg)
    m = getattr(importlib.import_module("models." + kind), kind)(cfg, d, dev)
    np.testing.assert_array_equal(m.user_embedding.weight.detach().shape, golden_tiny["lg_user_w0"].shape)
    _load_weights(m, golden_tiny["lg_user_w0"], golden_tiny["lg_item_w0"])
    m.to(dev)
    b = torch.from_numpy(golden_next["batch"].copy()).to(dev)
    losses = m(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
    total = losses[0]
    for l in losses[1:]:
        total = total + l
    total.backward()
    p = "%s_%s" % (kind.lower(), enc.lower())
    np.testing.assert_allclose([l.item() for l in losses], golden_next[p + "_loss"], rtol=RTOL)
    _close(m.user_embedding.weight.grad.cpu().numpy(), golden_next[p + "_gu"], rtol=1e-4)
    _close(m.item_embedding.weight.grad.cpu().numpy(), golden_next[p + "_gi"], rtol=1e-4)
    # evaluation surface: aggregate() is LightGCN's (or the ego tables for MF)
    fu, fi = m.final_embeddings()
    if enc == "LightGCN":
        _close(fu.cpu().numpy(), golden_tiny["lg_fu0"]); _close(fi.cpu().numpy(), golden_tiny["lg_fi0"])
    else:
        np.testing.assert_array_equal(fu.cpu().numpy(), golden_tiny["lg_user_w0"])


def test_lightcscf_margin_branch_vs_reference(dev, golden_dirs, golden_tiny, golden_next):
    """Trained-like weights and a 0.05 margin: the relu branch is live for a large share of the pairs."""
    from models.LightCSCF import LightCSCF
    cfg = _cfg("LightCSCF", batch_size=256, encoder="LightGCN", **dict(NEXT_CFG["LightCSCF"], lambda_margin=0.05))
    d = _data(golden_dirs, cfg)
    m = LightCSCF(cfg, d, dev)
    _load_weights(m, golden_tiny["lg_user_wT"], golden_tiny["lg_item_wT"])
    m.to(dev)
    b = torch.from_numpy(golden_next["batch"].copy()).to(dev)
    losses = m(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
    (losses[0] + losses[1]).backward()
    np.testing.assert_allclose([l.item() for l in losses], golden_next["lightcscf_margin_loss"], rtol=RTOL)
    _close(m.user_embedding.weight.grad.cpu().numpy(), golden_next["lightcscf_margin_gu"], rtol=1e-4)
    _close(m.item_embedding.weight.grad.cpu().numpy(), golden_next["lightcscf_margin_gi"], rtol=1e-4)


def test_sgl_vs_reference(dev, golden_dirs, golden_tiny, golden_next, monkeypatch):
    """SGL.py:60-89 with the reference's kept-edge draws injected into random.sample: the sub-graph adjacencies are
    bit-identical to tools.create_adj_mat's, losses and gradients equal the reference's."""
    from models.SGL import SGL, Trainer
    import utility.utility_function.tools as tools
    cfg = _cfg("SGL", batch_size=256, **NEXT_CFG["SGL"])
    d = _data(golden_dirs, cfg)
    keeps = [golden_next["sgl_keep0"].tolist(), golden_next["sgl_keep1"].tolist()]
    calls = []

    def fake_sample(population, k):
        j = len(calls)
        calls.append(k)
        assert k == len(keeps[j]) and len(population) == d.user_item_net.count_nonzero()
        return keeps[j]
    monkeypatch.setattr(random, "sample", fake_sample)
    subs = [tools.convert_sp_mat_to_sp_tensor(tools.create_adj_mat(d.user_item_net, "ed", 0.1)).to(dev) for _ in range(2)]
    monkeypatch.undo()
    for j, g in enumerate(subs):
        np.testing.assert_array_equal(g.csr.indptr.cpu().numpy(), golden_next["sgl_sub%d_indptr" % j])
        np.testing.assert_array_equal(g.csr.indices.cpu().numpy(), golden_next["sgl_sub%d_indices" % j])
        np.testing.assert_array_equal(g.csr.data.cpu().numpy().view(np.uint32), golden_next["sgl_sub%d_data" % j].view(np.uint32))
    m = SGL(cfg, d, dev)
    _load_weights(m, golden_tiny["lg_user_w0"], golden_tiny["lg_item_w0"])
    m.to(dev)
    b = torch.from_numpy(golden_next["batch"].copy()).to(dev)
    losses = m(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous(), subs[0], subs[1])
    (losses[0] + losses[1] + losses[2]).backward()
    np.testing.assert_allclose([l.item() for l in losses], golden_next["sgl_loss"], rtol=RTOL)
    _close(m.user_embedding.weight.grad.cpu().numpy(), golden_next["sgl_gu"], rtol=1e-4)
    _close(m.item_embedding.weight.grad.cpu().numpy(), golden_next["sgl_gi"], rtol=1e-4)
    # 'rw' (one sub-graph per layer) takes the per-layer product path; with the same graph on every layer it must
    # reproduce the fused K-layer propagation
    with torch.no_grad():
        fa = m.aggregate(subs[0])
        fb = m.aggregate([subs[0]] * 3)
    _close(fb[0].cpu().numpy(), fa[0].cpu().numpy()); _close(fb[1].cpu().numpy(), fa[1].cpu().numpy())
    assert Trainer is not None


@pytest.mark.parametrize("mode", ["parallel", "alternating"])
def test_egcf_vs_reference(dev, golden_dirs, golden_next, mode):
    """EGCF.py:45-112: the R graph equals the reference's bit for bit (float64 degrees, rounded to fp32 once), final
    embeddings, losses and the item-table gradient equal the reference's in both aggregate modes."""
    import utility.utility_function.tools as tools
    from models.EGCF import EGCF
    cfg = _cfg("EGCF", batch_size=256, mode=mode, ssl_lambda=0.1, temperature=0.1)
    d = _data(golden_dirs, cfg)
    tools.set_seed(2024)
    m = EGCF(cfg, d, dev)
    # same torch-generator draws as the reference: nn.Embedding's own init, then xavier on the item table only
    np.testing.assert_array_equal(m.item_embedding.weight.detach().numpy(), golden_next["egcf_item_w0"])
    R = m.user_Graph.csr.to_scipy()[:d.num_users, d.num_users:].tocoo()
    order = np.lexsort((R.col, R.row))
    np.testing.assert_array_equal(np.stack([R.row[order], R.col[order]]), golden_next["egcf_R_index"])
    np.testing.assert_array_equal(R.data[order].view(np.uint32), golden_next["egcf_R_value"].view(np.uint32))
    m.to(dev)
    fu, fi = m.final_embeddings()
    _close(fu.cpu().numpy(), golden_next["egcf_%s_fu" % mode]); _close(fi.cpu().numpy(), golden_next["egcf_%s_fi" % mode])
    b = torch.from_numpy(golden_next["batch"].copy()).to(dev)
    losses = m(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
    (losses[0] + losses[1] + losses[2]).backward()
    np.testing.assert_allclose([l.item() for l in losses], golden_next["egcf_%s_loss" % mode], rtol=RTOL)
    _close(m.item_embedding.weight.grad.cpu().numpy(), golden_next["egcf_%s_gi" % mode], rtol=1e-4)
    r = m.get_rating_for_test(torch.arange(7, device=dev))
    want = 1.0 / (1.0 + np.exp(-(golden_next["egcf_%s_fu" % mode][:7].astype(np.float64) @ golden_next["egcf_%s_fi" % mode].astype(np.float64).T)))
    np.testing.assert_allclose(r.cpu().numpy(), want, rtol=1e-5)


@pytest.mark.parametrize("kind", ["LightCCF", "DirectAU", "SGL", "EGCF"])
def test_trainers_run_end_to_end(dev, golden_dirs, kind, capsys):
    """Trainer(...).train() for two epochs on the tiny dataset: finite decreasing-or-equal losses, metrics in range,
    the same log lines as the reference."""
    import logging
    import utility.utility_function.tools as tools
    cfg = _cfg(kind, batch_size=256, training_epochs=2, interval=1, test_batch_size=37, **NEXT_CFG.get(kind, {}))
    if kind == "SGL":
        cfg.update(aug_type="ed", ssl_ratio="0.1")
    elif kind == "EGCF":
        cfg.update(mode="alternating", top_K="[10, 20]")
    else:
        cfg["encoder"] = "LightGCN"
    tools.set_seed(2024)
    d = _data(golden_dirs, cfg)
    logger = logging.getLogger("idgrec-test-" + kind)
    records = []

    class H(logging.Handler):
        def emit(self, r):
            records.append(r.getMessage())
    logger.addHandler(H())
    logger.setLevel(logging.INFO)
    tr = importlib.import_module("models." + kind).Trainer(None, cfg, d, dev, logger)
    w0 = tr.model.item_embedding.weight.detach().clone()
    tr.train()
    assert any("training loss" in r for r in records) and any("Test recall" in r for r in records)
    w1 = tr.model.item_embedding.weight.detach().cpu()
    assert torch.isfinite(w1).all() and not torch.equal(w1, w0.cpu())


# ------------------------------------------------------------------------------------------------
# evaluator: activity groups and a wider top-K list
# ------------------------------------------------------------------------------------------------
def test_sparsity_test_and_top40_vs_reference(dev, golden_dirs, golden_tiny, golden_next):
    """batch_test.sparsity_test (batch_test.py:110-170) and Test() with top_K = [20, 40] equal the reference's
    to 4 decimals; the per-group ids equal the exact-rank oracle bit for bit."""
    from models.LightGCN import LightGCN
    import utility.utility_train.batch_test as batch_test
    cfg = _cfg("LightGCN", test_batch_size=37, top_K="[20, 40]", sparsity_test=1)
    d = _data(golden_dirs, cfg)
    m = LightGCN(cfg, d, dev)
    _load_weights(m, golden_tiny["lg_user_wT"], golden_tiny["lg_item_wT"])
    m.to(dev)
    res = batch_test.sparsity_test(d, m, dev, cfg)
    assert len(res) == golden_next["sparsity_recall"].shape[0]
    for j, r in enumerate(res):
        for k in ("recall", "precision", "ndcg"):
            np.testing.assert_allclose(r[k], golden_next["sparsity_" + k][j], rtol=0, atol=5e-5, err_msg="group %d %s" % (j, k))
    first, best = batch_test.general_test(d, m, dev, cfg, 0, {"count": 0, "epoch": 0, "recall": [0.0, 0.0], "ndcg": [0.0, 0.0], "stop": 0})
    np.testing.assert_allclose(first["recall"], golden_next["sparsity_recall"][0], atol=5e-5)
    cfg0 = dict(cfg, sparsity_test="0")
    full = batch_test.Test(d, m, dev, cfg0)
    for k in ("recall", "precision", "ndcg"):
        np.testing.assert_allclose(full[k], golden_next["test2040_" + k], rtol=0, atol=5e-5, err_msg=k)
    # T0 on one group: K = 40 ids against the exact-rank oracle
    from idgrec import ops
    cache = d.device_cache(dev)
    fu, fi = m.final_embeddings()
    users = np.asarray(d.split_test_dict[1], dtype=np.int64)
    ids = ops.eval_topk(fu, fi, torch.from_numpy(users).to(dev), cache["mask_indptr"], cache["mask_indices"], 40)
    od = O.load_dataset(golden_dirs["tiny"])
    ref_ids, _ = O.topk_exact(fu.cpu().numpy(), fi.cpu().numpy(), users, od.user_item_net.indptr, od.user_item_net.indices, 40)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_ids)


# ------------------------------------------------------------------------------------------------
# whole-step CUDA-graph replay of the autograd-path models
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["LightCCF", "DirectAU", "NGCF", "EGCF"])
def test_graphed_step_equals_eager_loop(dev, golden_dirs, golden_tiny, kind):
    """GraphedStep (forward -> backward -> capturable Adam replayed from a CUDA graph, captured per batch size, warm-up
    undone) follows the same trajectory as the reference's eager loop (trainer.py:40-56) over 5 batches incl. a short one."""
    from idgrec.graphed import GraphedStep
    over = dict(NEXT_CFG.get(kind, {}))
    if kind == "NGCF":
        over = {"mess_drop_prob": "[0.0, 0.0, 0.0]"}      # dropout off: both runs must be deterministic
    elif kind == "EGCF":
        over = {"mode": "parallel"}
    else:
        over["encoder"] = "LightGCN"
    cfg = _cfg(kind, batch_size=256, **over)
    d = _data(golden_dirs, cfg)
    s0 = golden_tiny["sample_ep0"][golden_tiny["perm_ep0"]]
    batches = [torch.from_numpy(s0[a:b].copy()).to(dev) for a, b in ((0, 256), (256, 512), (512, 768), (768, 868), (868, 1124))]
    finals, sums = [], []
    for mode in ("eager", "graph"):
        import utility.utility_function.tools as tools
        tools.set_seed(2024)
        m = getattr(importlib.import_module("models." + kind), kind)(cfg, d, dev)
        m.to(dev)
        if mode == "eager":
            opt = torch.optim.Adam(m.parameters(), lr=1e-3)
            acc = None
            for b in batches:
                ll = m(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
                st = torch.stack([l.reshape(()) for l in ll])
                opt.zero_grad()
                st.sum().backward()
                opt.step()
                acc = st.detach() if acc is None else acc + st.detach()
            sums.append(acc.cpu().numpy())
        else:
            gs = GraphedStep(m, 1e-3, 256)
            for b in batches:
                gs.step(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
            assert sorted(gs.graphs) == [100, 256] and gs.replays == 5
            sums.append(np.asarray(gs.pop_epoch_losses()))
            assert gs.pop_epoch_losses() == [0.0] * len(sums[-1])
        finals.append([p.detach().cpu().numpy().copy() for p in m.parameters()])
    np.testing.assert_allclose(sums[1], sums[0], rtol=2e-5)
    for a, b in zip(finals[1], finals[0]):
        _close(a, b, rtol=2e-5)


# ------------------------------------------------------------------------------------------------
# the same models on the fused CUDA-graph step (idgrec.engine.FusedTrainer)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["LightCCF", "LightCSCF", "SCCF", "DirectAU"])
def test_pair_models_fused_step_vs_reference(dev, golden_dirs, golden_tiny, golden_next, kind):
    """One fused step (row-restricted propagation, BPR / reg selected by upstream weights, batch x batch loss on tensor cores,
    deterministic scatter into the shared backward propagation): losses and the full gradient equal the reference's."""
    for restrict in (1, 0):
        cfg = _cfg(kind, batch_size=256, encoder="LightGCN", cuda_graph=0, restrict_rows=restrict, fuse_adam=0, **NEXT_CFG[kind])
        d = _data(golden_dirs, cfg)
        m = getattr(importlib.import_module("models." + kind), kind)(cfg, d, dev)
        _load_weights(m, golden_tiny["lg_user_w0"], golden_tiny["lg_item_w0"])
        m.to(dev)
        ft = m.fused_trainer(1e-3, 256)
        assert ft is not None and ft.kind == kind
        b = torch.from_numpy(golden_next["batch"].copy()).to(dev)
        loss = ft.step(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous(), apply_adam=False)
        p = "%s_lightgcn" % kind.lower()
        np.testing.assert_allclose(loss.cpu().numpy(), golden_next[p + "_loss"], rtol=RTOL)
        _close(ft.gE0[:d.num_users].cpu().numpy(), golden_next[p + "_gu"], rtol=1e-4)
        _close(ft.gE0[d.num_users:].cpu().numpy(), golden_next[p + "_gi"], rtol=1e-4)
        assert float(ft.G.abs().max()) == 0.0        # the gradient scratch table is clean again after the step
    # MF encoder keeps the autograd path
    cfg = _cfg(kind, batch_size=256, encoder="MF", **NEXT_CFG[kind])
    m = getattr(importlib.import_module("models." + kind), kind)(cfg, _data(golden_dirs, cfg), dev)
    m.to(dev)
    assert m.fused_trainer(1e-3, 256) is None


@pytest.mark.parametrize("kind", ["LightCCF", "SCCF", "DirectAU"])
def test_pair_models_fused_graph_trajectory(dev, golden_dirs, golden_tiny, kind):
    """Five batches (one short) through the captured fused step with Adam in the last backward epilogue follow the eager
    autograd + torch.optim.Adam loop of the reference (trainer.py:40-56)."""
    s0 = golden_tiny["sample_ep0"][golden_tiny["perm_ep0"]]
    batches = [torch.from_numpy(s0[a:b].copy()).to(dev) for a, b in ((0, 256), (256, 512), (512, 768), (768, 868), (868, 1124))]
    finals, sums = [], []
    for mode in ("eager", "fused"):
        cfg = _cfg(kind, batch_size=256, encoder="LightGCN", **NEXT_CFG[kind])
        d = _data(golden_dirs, cfg)
        m = getattr(importlib.import_module("models." + kind), kind)(cfg, d, dev)
        _load_weights(m, golden_tiny["lg_user_w0"], golden_tiny["lg_item_w0"])
        m.to(dev)
        if mode == "eager":
            opt = torch.optim.Adam(m.parameters(), lr=1e-3)
            acc = None
            for b in batches:
                ll = m(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
                st = torch.stack([l.reshape(()) for l in ll])
                opt.zero_grad()
                st.sum().backward()
                opt.step()
                acc = st.detach() if acc is None else acc + st.detach()
            sums.append(acc.cpu().numpy())
        else:
            ft = m.fused_trainer(1e-3, 256)
            for b in batches:
                ft.step(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
            sums.append(np.asarray(ft.pop_epoch_losses()))
        finals.append([m.user_embedding.weight.detach().cpu().numpy().copy(), m.item_embedding.weight.detach().cpu().numpy().copy()])
    np.testing.assert_allclose(sums[1], sums[0], rtol=2e-5)
    for a, b in zip(finals[1], finals[0]):
        _close(a, b, rtol=2e-5)


# ------------------------------------------------------------------------------------------------
# NGCF on its fused step (idgrec.engine_ngcf.NgcfFusedTrainer)
# ------------------------------------------------------------------------------------------------
def _ngcf(dev, golden_dirs, golden_tiny, **over):
    from models.NGCF import NGCF
    g = golden_tiny
    cfg = _cfg("NGCF", **over)
    d = _data(golden_dirs, cfg)
    m = NGCF(cfg, d, dev)
    _load_weights(m, g["ngcf_user_w0"], g["ngcf_item_w0"])
    with torch.no_grad():
        for l in range(3):
            for k in ("W_gcn", "b_gcn", "W_bi", "b_bi"):
                m.weight_dict["%s_%d" % (k, l)].copy_(torch.from_numpy(g["ngcf_%s_%d" % (k, l)]))
    m.to(dev)
    return m, d


def test_ngcf_fused_step_vs_reference(dev, golden_dirs, golden_tiny):
    """One fused NGCF step with the reference's dropout masks injected: losses, table gradients and all 12 dense-weight
    gradients equal the unmodified reference's (ngcf_* in tiny.npz)."""
    g = golden_tiny
    m, d = _ngcf(dev, golden_dirs, g)
    ft = m.fused_trainer(1e-4, 256)
    ft.injected_keep = [torch.from_numpy(x).to(dev) for x in g["ngcf_masks"]]
    b = torch.from_numpy(g["batch"].copy()).to(dev)
    loss = ft.step(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous(), apply_adam=False)
    np.testing.assert_allclose(loss.cpu().numpy(), g["ngcf_loss"], rtol=RTOL)
    _close(ft.gE0[:d.num_users].cpu().numpy(), g["ngcf_gu"], rtol=1e-4)
    _close(ft.gE0[d.num_users:].cpu().numpy(), g["ngcf_gi"], rtol=1e-4)
    for l in range(3):
        for k in ("W_gcn", "b_gcn", "W_bi", "b_bi"):
            name = "%s_%d" % (k, l)
            _close(ft.gw[name].cpu().numpy().reshape(g["ngcf_g_" + name].shape), g["ngcf_g_" + name], rtol=1e-4)
    assert float(ft.G.abs().max()) == 0.0 and float(ft.G64.abs().max()) == 0.0
    # the parameters are views of the flat buffer: evaluation sees what Adam updates
    assert m.weight_dict["W_bi_2"].data_ptr() == ft.w["W_bi_2"].data_ptr()


def test_ngcf_fused_step_equals_eager_loop(dev, golden_dirs, golden_tiny):
    """Five batches (one short) through the captured fused step follow the autograd ops + torch.optim.Adam loop (dropout off)."""
    s0 = golden_tiny["sample_ep0"][golden_tiny["perm_ep0"]]
    batches = [torch.from_numpy(s0[a:b].copy()).to(dev) for a, b in ((0, 256), (256, 512), (512, 768), (768, 868), (868, 1124))]
    finals, sums = [], []
    for mode in ("eager", "fused"):
        m, d = _ngcf(dev, golden_dirs, golden_tiny, mess_drop_prob="[0.0, 0.0, 0.0]", fused_step=int(mode == "fused"))
        if mode == "eager":
            assert m.fused_trainer(1e-3, 256) is None
            opt = torch.optim.Adam(m.parameters(), lr=1e-3)
            acc = None
            for b in batches:
                ll = m(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
                st = torch.stack([l.reshape(()) for l in ll])
                opt.zero_grad()
                st.sum().backward()
                opt.step()
                acc = st.detach() if acc is None else acc + st.detach()
            sums.append(acc.cpu().numpy())
        else:
            ft = m.fused_trainer(1e-3, 256)
            for b in batches:
                ft.step(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
            assert sorted(ft._graphs) == [100, 256]
            sums.append(np.asarray(ft.pop_epoch_losses()))
        finals.append([p.detach().cpu().numpy().copy() for p in m.parameters()])
    np.testing.assert_allclose(sums[1], sums[0], rtol=2e-5)
    for a, b in zip(finals[1], finals[0]):
        _close(a, b, rtol=2e-5)
