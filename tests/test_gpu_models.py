"""Model-level parity on the GPU: the host mirror (models/*.py, utility_train) driven exactly like
the reference drives its own classes, checked against golden outputs of the unmodified reference
(tests/golden/tiny.npz) and against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import REPO
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _cfg(name, **over):
    import utility.utility_function.tools as tools
    c = tools.read_configuration(os.path.join(REPO, "id-grec_b200", "configure", name + ".txt"), name)
    c.update(dataset="tiny", **{k: str(v) for k, v in over.items()})
    return c


def _data(golden_dirs, cfg):
    from utility.utility_data.data_loader import Data
    return Data(golden_dirs["tiny"], cfg)


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * max(np.abs(b).max(), 1e-30))


def _load_weights(model, uw, iw):
    with torch.no_grad():
        model.user_embedding.weight.copy_(torch.from_numpy(uw))
        model.item_embedding.weight.copy_(torch.from_numpy(iw))


def test_model_init_matches_reference_rng_order(dev, golden_dirs, golden_tiny):
    """set_seed(2024) then user table, item table xavier draws (LightGCN.py:27-28)."""
    import utility.utility_function.tools as tools
    from models.LightGCN import LightGCN
    cfg = _cfg("LightGCN")
    d = _data(golden_dirs, cfg)
    tools.set_seed(2024)
    m = LightGCN(cfg, d, dev)
    np.testing.assert_array_equal(m.user_embedding.weight.detach().numpy(), golden_tiny["lg_user_w0"])
    np.testing.assert_array_equal(m.item_embedding.weight.detach().numpy(), golden_tiny["lg_item_w0"])


@pytest.mark.parametrize("path", ["autograd", "fused", "fused_graph", "fused_dense", "fused_closure"])
def test_lightgcn_two_steps_vs_reference(dev, golden_dirs, golden_tiny, path):
    """Reference trainer loop (trainer.py:40-56) for two batches of 256: losses, gradients and the
    Adam-updated tables equal the unmodified reference's."""
    from models.LightGCN import LightGCN
    g = golden_tiny
    # "fused": Adam-fused epilogue, eager; "fused_graph": same, replayed from a CUDA graph; "fused_dense": no row
    # restriction and a separate Adam kernel (so the gradient table can be compared as well)
    cfg = _cfg("LightGCN", batch_size=256, cuda_graph=int(path == "fused_graph"), restrict_rows=int(path != "fused_dense"),
               fuse_adam=int(path != "fused_dense"), closure_restrict=int(path == "fused_closure"))
    d = _data(golden_dirs, cfg)
    m = LightGCN(cfg, d, dev)
    _load_weights(m, g["lg_user_w0"], g["lg_item_w0"])
    m.to(dev)
    fu, fi = m.final_embeddings()
    _close(fu.cpu().numpy(), g["lg_fu0"]); _close(fi.cpu().numpy(), g["lg_fi0"])
    s0 = g["sample_ep0"][g["perm_ep0"]]
    opt = torch.optim.Adam(m.parameters(), lr=1e-3) if path == "autograd" else None
    ft = None if path == "autograd" else m.fused_trainer(1e-3, 256)
    for st in range(2):
        b = torch.from_numpy(s0[st * 256:(st + 1) * 256].copy()).to(dev)
        if path == "autograd":
            losses = m(b[:, 0], b[:, 1], b[:, 2])
            total = losses[0] + losses[1]
            opt.zero_grad()
            total.backward()
            np.testing.assert_allclose([l.item() for l in losses], g["lg_loss_s%d" % st], rtol=RTOL)
            _close(m.user_embedding.weight.grad.cpu().numpy(), g["lg_gu_s%d" % st])
            _close(m.item_embedding.weight.grad.cpu().numpy(), g["lg_gi_s%d" % st])
            opt.step()
        else:
            loss = ft.step(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
            np.testing.assert_allclose(loss.cpu().numpy(), g["lg_loss_s%d" % st], rtol=RTOL)
            if not ft.fuse_adam:
                _close(ft.gE0[:d.num_users].cpu().numpy(), g["lg_gu_s%d" % st])
                _close(ft.gE0[d.num_users:].cpu().numpy(), g["lg_gi_s%d" % st])
            else:
                assert float(ft.regc.abs().max()) == 0.0 and float(ft.G.abs().max()) == 0.0
        # Adam's first steps move every weight by ~lr*g/(|g|+1e-8): entries with |g| ~ 1e-8 amplify the
        # 1e-5 gradient tolerance, so the updated tables are compared at 1e-5 of the table scale
        _close(m.user_embedding.weight.detach().cpu().numpy(), g["lg_user_w_s%d" % st])
        _close(m.item_embedding.weight.detach().cpu().numpy(), g["lg_item_w_s%d" % st])
    if ft is not None:
        acc = ft.pop_epoch_losses()
        np.testing.assert_allclose(acc, g["lg_loss_s0"] + g["lg_loss_s1"], rtol=RTOL)


@pytest.mark.parametrize("tag,wu,wi", [("lg", "lg_user_w_s1", "lg_item_w_s1"), ("lgT", "lg_user_wT", "lg_item_wT")])
def test_Test_metrics_vs_reference(dev, golden_dirs, golden_tiny, tag, wu, wi):
    """T2: utility_train.batch_test.Test() equals the reference's own Test() to 4 decimals."""
    from models.LightGCN import LightGCN
    import utility.utility_train.batch_test as batch_test
    g = golden_tiny
    cfg = _cfg("LightGCN", test_batch_size=50)
    d = _data(golden_dirs, cfg)
    m = LightGCN(cfg, d, dev)
    _load_weights(m, g[wu], g[wi])
    m.to(dev)
    res = batch_test.Test(d, m, dev, cfg)
    for k in ("recall", "precision", "ndcg"):
        np.testing.assert_allclose(res[k], g[tag + "_test_" + k], rtol=0, atol=5e-5, err_msg=k)
    assert res["hit"].tolist() == [0.0, 0.0]
    # T0: ids bit-exact against the exact-rank oracle on the same embeddings
    ids, users = batch_test.rank_all(d, m, dev, 20)
    fu, fi = (t.cpu().numpy() for t in m.final_embeddings())
    od = O.load_dataset(golden_dirs["tiny"])
    ref_ids, _ = O.topk_exact(fu, fi, users.cpu().numpy(), od.user_item_net.indptr, od.user_item_net.indices, 20)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_ids)
    # get_rating_for_test keeps the reference's dense [b, I] contract
    if tag == "lgT":
        r = m.get_rating_for_test(torch.from_numpy(g["test_users"][:50]).to(dev))
        np.testing.assert_allclose(r.cpu().numpy(), g["lgT_rating50"], rtol=1e-5)


@pytest.mark.parametrize("kind,prefix", [("SimGCL", "simgcl"), ("XSimGCL", "xsimgcl"), ("XSimGCL", "xsimgcl2")])
@pytest.mark.parametrize("path", ["autograd", "fused"])
def test_contrastive_models_vs_reference(dev, golden_dirs, golden_tiny, kind, prefix, path):
    """SimGCL / XSimGCL forward + backward with the reference's own noise draws injected."""
    import importlib
    g = golden_tiny
    cfg = _cfg(kind, batch_size=256, **({"cl_layer": 2} if prefix == "xsimgcl2" else {}))
    d = _data(golden_dirs, cfg)
    m = getattr(importlib.import_module("models." + kind), kind)(cfg, d, dev)
    _load_weights(m, g[prefix + "_user_w0"], g[prefix + "_item_w0"])
    m.to(dev)
    noise = torch.from_numpy(g[prefix + "_noise"]).to(dev)
    b = torch.from_numpy(g["batch"].copy()).to(dev)
    u, p, n = (b[:, k].contiguous() for k in range(3))
    if path == "autograd":
        losses = m(u, p, n, (noise[:3].contiguous(), noise[3:6].contiguous())) if kind == "SimGCL" else m(u, p, n, noise[:3].contiguous())
        (losses[0] + losses[1] + losses[2]).backward()
        got = [l.item() for l in losses]
        gu, gi = m.user_embedding.weight.grad.cpu().numpy(), m.item_embedding.weight.grad.cpu().numpy()
    else:
        ft = m.fused_trainer(1e-3, 256)
        ft.injected_noise = [noise[:3].contiguous(), noise[3:6].contiguous()] if kind == "SimGCL" else [noise[:3].contiguous()]
        got = ft.step(u, p, n, apply_adam=False).cpu().numpy()
        gu, gi = ft.gE0[:d.num_users].cpu().numpy(), ft.gE0[d.num_users:].cpu().numpy()
        # the gradient scratch tables are clean again after the step
        assert float(ft.G.abs().max()) == 0.0
        if ft.Gcl is not None:
            assert float(ft.Gcl.abs().max()) == 0.0
    np.testing.assert_allclose(got, g[prefix + "_loss"], rtol=RTOL)
    _close(gu, g[prefix + "_gu"], rtol=1e-4)
    _close(gi, g[prefix + "_gi"], rtol=1e-4)
    fu, fi = m.final_embeddings()
    _close(fu.cpu().numpy(), g[prefix + "_fu"]); _close(fi.cpu().numpy(), g[prefix + "_fi"])


def test_ngcf_vs_reference(dev, golden_dirs, golden_tiny):
    """NGCF forward + backward with the reference's dropout masks injected (ngcf_* in tiny.npz), then the
    256-d full-ranking evaluation against the exact-rank oracle."""
    from models.NGCF import NGCF
    import utility.utility_train.batch_test as batch_test
    g = golden_tiny
    cfg = _cfg("NGCF")
    d = _data(golden_dirs, cfg)
    m = NGCF(cfg, d, dev)
    _load_weights(m, g["ngcf_user_w0"], g["ngcf_item_w0"])
    with torch.no_grad():
        for l in range(3):
            for k in ("W_gcn", "b_gcn", "W_bi", "b_bi"):
                m.weight_dict["%s_%d" % (k, l)].copy_(torch.from_numpy(g["ngcf_%s_%d" % (k, l)]))
    m.to(dev)
    masks = [torch.from_numpy(x).to(dev) for x in g["ngcf_masks"]]
    b = torch.from_numpy(g["batch"].copy()).to(dev)
    losses = m(b[:, 0], b[:, 1], b[:, 2], keep_masks=masks)
    (losses[0] + losses[1]).backward()
    np.testing.assert_allclose([l.item() for l in losses], g["ngcf_loss"], rtol=RTOL)
    _close(m.user_embedding.weight.grad.cpu().numpy(), g["ngcf_gu"], rtol=1e-4)
    _close(m.item_embedding.weight.grad.cpu().numpy(), g["ngcf_gi"], rtol=1e-4)
    for l in range(3):
        for k in ("W_gcn", "b_gcn", "W_bi", "b_bi"):
            _close(m.weight_dict["%s_%d" % (k, l)].grad.cpu().numpy(), g["ngcf_g_%s_%d" % (k, l)], rtol=1e-4)
    # 256-d ranking through the fused evaluator (dropout is live at eval in the reference: fix the masks)
    with torch.no_grad():
        fu, fi = m.aggregate(masks)
    od = O.load_dataset(golden_dirs["tiny"])
    users = np.array(list(od.test_dict.keys()), dtype=np.int64)
    cache = d.device_cache(dev)
    from idgrec import ops
    ids = ops.eval_topk(fu.contiguous(), fi.contiguous(), cache["test_users"], cache["mask_indptr"], cache["mask_indices"], 20)
    ref_ids, _ = O.topk_exact(fu.cpu().numpy(), fi.cpu().numpy(), users, od.user_item_net.indptr, od.user_item_net.indices, 20)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_ids)
    res = batch_test.Test(d, m, dev, cfg)
    assert 0.0 <= res["recall"][1] <= 1.0


@pytest.mark.parametrize("kind", ["SimGCL", "XSimGCL"])
def test_contrastive_graph_replay_equals_eager(dev, golden_dirs, golden_tiny, kind):
    """The CUDA-graph path of the contrastive models (device-side unique, device row counts, Adam-fused last
    layer) gives the same tables as the eager path on the same batches and noise."""
    import importlib
    g = golden_tiny
    prefix = kind.lower()
    noise = torch.from_numpy(g[prefix + "_noise"]).to(dev)
    inj = [noise[:3].contiguous(), noise[3:6].contiguous()] if kind == "SimGCL" else [noise[:3].contiguous()]
    s0 = g["sample_ep0"][g["perm_ep0"]]
    tables, losses = [], []
    for graph in (0, 1):
        cfg = _cfg(kind, batch_size=256, cuda_graph=graph)
        d = _data(golden_dirs, cfg)
        m = getattr(importlib.import_module("models." + kind), kind)(cfg, d, dev)
        _load_weights(m, g[prefix + "_user_w0"], g[prefix + "_item_w0"])
        m.to(dev)
        ft = m.fused_trainer(1e-3, 256)
        ft.injected_noise = inj
        for st in range(3):
            b = torch.from_numpy(s0[st * 256:(st + 1) * 256].copy()).to(dev)
            ft.step(b[:, 0].contiguous(), b[:, 1].contiguous(), b[:, 2].contiguous())
        losses.append(ft.pop_epoch_losses())
        tables.append(m._table.clone())
    np.testing.assert_allclose(losses[0], losses[1], rtol=1e-6)
    assert torch.equal(tables[0], tables[1])


def test_mfbpr_vs_reference(dev, golden_dirs, golden_tiny):
    from models.MFBPR import MFBPR
    g = golden_tiny
    cfg = _cfg("MFBPR")
    d = _data(golden_dirs, cfg)
    m = MFBPR(cfg, d, dev)
    _load_weights(m, g["lg_user_w0"], g["lg_item_w0"])
    m.to(dev)
    b = torch.from_numpy(g["batch"].copy()).to(dev)
    losses = m(b[:, 0], b[:, 1], b[:, 2])
    (losses[0] + losses[1]).backward()
    np.testing.assert_allclose([l.item() for l in losses], g["mf_loss"], rtol=RTOL)
    _close(m.user_embedding.weight.grad.cpu().numpy(), g["mf_gu"])


def test_functional_losses_api(dev, golden_tiny):
    """utility_function.losses keeps the reference's tensor-in/tensor-out functions (losses.py:4-35)."""
    import utility.utility_function.losses as losses
    g = golden_tiny
    a, b, c = (torch.from_numpy(g[k]).to(dev).requires_grad_(True) for k in ("fn_a", "fn_b", "fn_c"))
    bpr, reg, nce = losses.get_bpr_loss(a, b, c), losses.get_reg_loss(a, b, c), losses.get_InfoNCE_loss(a, b, 0.2)
    np.testing.assert_allclose([bpr.item(), reg.item(), nce.item()], [float(g["fn_bpr"]), float(g["fn_reg"]), float(g["fn_nce"])], rtol=RTOL)
    (bpr + reg + nce).backward()
    ar, br, cr = (torch.from_numpy(g[k]).requires_grad_(True) for k in ("fn_a", "fn_b", "fn_c"))
    (O.bpr_loss(ar, br, cr) + O.reg_loss(ar, br, cr) + O.infonce_loss(ar, br, 0.2)).backward()
    for x, y in ((a, ar), (b, br), (c, cr)):
        _close(x.grad.cpu().numpy(), y.grad.numpy(), rtol=5e-5)


def test_universal_trainer_end_to_end(dev, golden_dirs, tmp_path):
    """Two epochs through utility_train.trainer.universal_trainer: the epoch's batches come from the
    reference's sampler/shuffle stream and the logged losses equal the oracle model trained on the
    same batches."""
    import logging
    import utility.utility_function.tools as tools
    import utility.utility_train.trainer as trainer
    from models.LightGCN import LightGCN
    cfg = _cfg("LightGCN", batch_size=512, training_epochs=2, interval=1)
    d = _data(golden_dirs, cfg)
    tools.set_seed(2024)
    m = LightGCN(cfg, d, dev)
    w0 = (m.user_embedding.weight.detach().numpy().copy(), m.item_embedding.weight.detach().numpy().copy())
    logger = logging.getLogger("idg_test")
    logger.setLevel(logging.INFO)
    logf = tmp_path / "log.txt"
    h = logging.FileHandler(str(logf))
    logger.addHandler(h)
    trainer.universal_trainer(m, None, cfg, d, dev, logger)
    h.flush()
    lines = open(str(logf)).read().strip().split("\n")
    assert sum("Training time" in l for l in lines) == 2 and sum("Test recall" in l for l in lines) == 2
    # oracle on the same stream
    od = O.load_dataset(golden_dirs["tiny"])
    ip, ix, dt, _ = O.norm_adjacency(od.user_item_net)
    om = O.OracleModel("LightGCN", O.csr_to_torch_coo(ip, ix, dt, od.num_nodes), w0[0], w0[1])
    np.random.seed(2024)
    for ep in range(2):
        s = O.sample_negatives_bulk(od)
        s = s[O.shuffle_indices(len(s))]
        tot = np.zeros(2)
        for a, b in O.mini_batches(len(s), 512):
            tot += om.step(s[a:b, 0], s[a:b, 1], s[a:b, 2]).losses
        nb = len(s) // 512 + 1
        want = str(round(tot.sum() / nb, 6))
        line = [l for l in lines if "Training time" in l][ep]
        got = line.split("training loss: ")[1].split(" = ")[0]
        assert abs(float(got) - float(want)) <= 2e-6, (got, want)
    np.testing.assert_allclose(m.user_embedding.weight.detach().cpu().numpy(), om.user_w.detach().numpy(), rtol=1e-4, atol=1e-7)
