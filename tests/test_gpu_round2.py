"""Round-2 parity additions (VERDICT r1 "close the parity gaps" + ADVICE): InfoNCE at the production size, the T1
ranking report against the reference-faithful fp32-sigmoid ranking, evaluation outside the tiled shapes (any
embedding width / top_K the reference accepts), the evaluator on a foreign module that only offers
``get_rating_for_test``, the bounded peer barrier, and the reference's own ``main.py`` text driving the package."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "id-grec_b200")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from idgrec import _lib
    _lib.lib()
    return torch.device("cuda:0")


def _assert_close(a, b, rtol, scale=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.abs(b).max() if scale is None else scale
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * max(s, 1e-30))


def _rand_net(U, I, E, seed):
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    u = rng.integers(0, U, E)
    i = (rng.zipf(1.3, E) - 1) % I
    key = np.unique(u.astype(np.int64) * I + i)
    net = sp.csr_matrix((np.ones(len(key)), (key // I, key % I)), shape=(U, I))
    net.sort_indices()
    return net


def _mask(net, dev):
    return torch.from_numpy(net.indptr.astype(np.int32)).to(dev), torch.from_numpy(net.indices.astype(np.int32)).to(dev)


# ---------------------------------------------------------------- a10 at the production size
@pytest.mark.parametrize("n,tau", [(2048, 0.2), (2048, 0.15), (1923, 0.2)])
def test_infonce_production_size(dev, n, tau):
    """n = 2,048 rows (a full SimGCL batch without duplicates; 1,923 = the measured mean of unique users)
    on the tcgen05 path.  Loss to 1e-5; gradients to 1e-5 of the largest gradient entry against the fp64 autograd of
    the same formula (the fp32 CPU autograd the oracle runs is itself only ~3e-5 from fp64 at this size, which is why
    the small-n test states 5e-5 against it)."""
    from idgrec import ops
    N, d = n + 100, 64
    gen = torch.Generator().manual_seed(n + int(tau * 100))
    V1 = torch.randn(N, d, generator=gen) * 0.1
    V2 = V1 + 0.03 * torch.randn(N, d, generator=gen)
    idx = torch.sort(torch.randperm(N, generator=gen)[:n]).values
    r1, r2 = V1.double().requires_grad_(True), V2.double().requires_grad_(True)
    ref = O.infonce_loss(r1[idx], r2[idx], tau)
    ref.backward()
    ref32 = O.infonce_loss(V1[idx], V2[idx], tau)
    g1, g2 = V1.to(dev).requires_grad_(True), V2.to(dev).requires_grad_(True)
    out = ops.infonce_rows(g1, g2, idx.to(dev), tau)
    np.testing.assert_allclose(out.item(), ref.item(), rtol=1e-5)
    np.testing.assert_allclose(out.item(), ref32.item(), rtol=1e-5)
    out.backward()
    _assert_close(g1.grad.cpu().numpy(), r1.grad.numpy(), rtol=1e-5)
    _assert_close(g2.grad.cpu().numpy(), r2.grad.numpy(), rtol=1e-5)


# ---------------------------------------------------------------- T1 ranking report
@pytest.mark.parametrize("scale,label", [(0.4, "trained-like"), (0.6, "saturated"), (0.03, "xavier-like")])
def test_t1_agreement_with_reference_faithful_ranking(dev, scale, label):
    """SURVEY section 7 tier T1: the device ranking (exact fp64 order, ties by id) against the reference's own arithmetic
    (fp32 matmul -> fp32 sigmoid -> -1 mask -> sort).  Every position where the two differ must be explained by the
    reference's rounding: the two items involved are an exact fp32-sigmoid tie or their fp32 ratings are ordered the other way
    by less than the fp32 matmul+sigmoid error bound.  The agreement rate is printed (-s) as the T1 report."""
    from idgrec import ops
    U, I, d, K = 600, 4000, 64, 20
    net = _rand_net(U, I, 25000, 5)
    gen = torch.Generator().manual_seed(11)
    Fu = (torch.randn(U, d, generator=gen) * scale).numpy()
    Fi = (torch.randn(I, d, generator=gen) * scale).numpy()
    users = np.arange(U, dtype=np.int64)
    mp, mi = _mask(net, dev)
    ids = ops.eval_topk(torch.from_numpy(Fu).to(dev), torch.from_numpy(Fi).to(dev), torch.from_numpy(users).to(dev), mp, mi, K).cpu().numpy()
    ref_ids, rating = O.topk_reference_faithful(Fu, Fi, users, net.indptr, net.indices, K)
    exact = O.scores_fp64_sequential(Fu, Fi)
    same = ids == ref_ids
    n_ties = n_near = 0
    for r, c in zip(*np.nonzero(~same)):
        a, b = ids[r, c], ref_ids[r, c]          # ours, reference's at this position
        ra, rb = rating[r, a], rating[r, b]
        if ra == rb:
            n_ties += 1                          # fp32 sigmoid tie: torch.topk's order is arbitrary there
            continue
        # fp32 near-tie: the reference's fp32 ratings invert the exact order; the exact scores must then be within
        # the fp32 error of a length-64 dot product (gamma_64 |u||i|) of each other
        bound = 2.0 * 64 * 2.0 ** -24 * np.linalg.norm(Fu[r]) * max(np.linalg.norm(Fi[a]), np.linalg.norm(Fi[b])) + 2.0 ** -22
        assert abs(exact[r, a] - exact[r, b]) <= bound, (label, r, c, a, b, exact[r, a], exact[r, b], ra, rb)
        n_near += 1
    # as sets the two top-K lists may only differ through items tied (or near-tied) at the K-th place: covered above
    agree = same.mean()
    print("T1 report [%s]: position agreement %.4f%% (%d / %d), sigmoid ties %d, fp32 near-ties %d"
          % (label, 100 * agree, same.sum(), same.size, n_ties, n_near))
    if label == "trained-like":
        assert agree > 0.99


# ---------------------------------------------------------------- evaluation outside the tiled shapes
@pytest.mark.parametrize("d,K", [(48, 20), (64, 60), (128, 20), (32, 20), (192, 30)])
def test_eval_generic_width_and_topk(dev, d, K):
    """Any embedding_size / top_K the reference would accept ranks to the same exact ids (eval.cu limits)."""
    from idgrec import ops
    U, I = 150, 900
    net = _rand_net(U, I, 6000, d + K)
    gen = torch.Generator().manual_seed(d * 1000 + K)
    Fu = (torch.randn(U, d, generator=gen) * 0.3).numpy()
    Fi = (torch.randn(I, d, generator=gen) * 0.3).numpy()
    Fi[40:60] = Fi[40]
    users = np.arange(U, dtype=np.int64)
    ref_ids, _ = O.topk_exact(Fu, Fi, users, net.indptr, net.indices, K)
    mp, mi = _mask(net, dev)
    ids = ops.eval_topk(torch.from_numpy(Fu).to(dev), torch.from_numpy(Fi).to(dev), torch.from_numpy(users).to(dev), mp, mi, K)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_ids)
    # metrics with cut-offs beyond 64 positions
    rng = np.random.default_rng(1)
    truth = [sorted(rng.choice(I, size=rng.integers(1, 30), replace=False).tolist()) for _ in range(U)]
    tptr = np.zeros(U + 1, np.int32)
    tptr[1:] = np.cumsum([len(t) for t in truth])
    tl = np.concatenate(truth).astype(np.int32)
    ks = sorted({min(10, K), K // 2, K})
    sums = ops.eval_metric_sums(ids, torch.from_numpy(users).to(dev), torch.from_numpy(tptr).to(dev), torch.from_numpy(tl).to(dev), ks).cpu().numpy()
    r = O.hit_matrix(ref_ids, truth)
    tlen = np.array([len(t) for t in truth], dtype=np.float64)
    for j, k in enumerate(ks):
        np.testing.assert_allclose(sums[j], O.metric_sums(r, tlen, k), rtol=1e-12)


def test_config_validation_before_training(dev, golden_dirs):
    """A configuration outside the kernels is rejected when the model is constructed, not inside the first Test()."""
    from utility.utility_data.data_loader import Data
    from models.SimGCL import SimGCL
    from models.LightGCN import LightGCN
    cfg = {"dataset_path": os.path.dirname(golden_dirs["tiny"]) + "/", "dataset": "tiny", "top_K": "[10, 20]", "embedding_size": "48",
           "batch_size": "256", "test_batch_size": "50", "learn_rate": "0.001", "reg_lambda": "0.0001", "GCN_layer": "3", "sparsity_test": "0",
           "ssl_lambda": "0.1", "epsilon": "0.1", "temperature": "0.2"}
    data = Data(golden_dirs["tiny"], cfg)
    with pytest.raises(ValueError, match="embedding_size = 48"):
        LightGCN(cfg, data, dev)
    cfg["embedding_size"] = "128"
    LightGCN(cfg, data, dev)                      # 128 is fine for LightGCN ...
    with pytest.raises(ValueError, match="InfoNCE"):
        SimGCL(cfg, data, dev)                    # ... but not for the InfoNCE kernels
    cfg["embedding_size"] = "64"
    cfg["top_K"] = "[1, 2, 3, 4, 5, 6, 7, 8, 9]"
    with pytest.raises(ValueError, match="at most 8"):
        LightGCN(cfg, data, dev)


def test_lightgcn_width_128_top50_end_to_end(dev, golden_dirs):
    """embedding_size = 128, top_K = [20, 50]: one fused step + Test() against the oracle (reference-legal shape that the
    round-1 evaluator rejected after the first epoch)."""
    import utility.utility_function.tools as tools
    import utility.utility_train.batch_test as batch_test
    from utility.utility_data.data_loader import Data
    from models.LightGCN import LightGCN
    cfg = {"dataset_path": os.path.dirname(golden_dirs["tiny"]) + "/", "dataset": "tiny", "top_K": "[20, 50]", "embedding_size": "128",
           "batch_size": "256", "test_batch_size": "50", "learn_rate": "0.001", "reg_lambda": "0.0001", "GCN_layer": "3", "sparsity_test": "0"}
    data = Data(golden_dirs["tiny"], cfg)
    tools.set_seed(2024)
    model = LightGCN(cfg, data, dev).to(dev)
    od = O.load_dataset(golden_dirs["tiny"])
    ip, ix, dt, _ = O.norm_adjacency(od.user_item_net)
    A = O.csr_to_torch_coo(ip, ix, dt, od.num_nodes)
    om = O.OracleModel("LightGCN", A, model.user_embedding.weight.detach().cpu().numpy().copy(), model.item_embedding.weight.detach().cpu().numpy().copy())
    rng = np.random.default_rng(3)
    e = rng.integers(0, len(od.train_user), 256)
    bu, bp, bn = od.train_user[e], od.train_item[e], rng.integers(0, od.num_items, 256)
    ref = om.step(bu, bp, bn)
    ft = model.fused_trainer(1e-3, 256)
    loss = ft.step(*(torch.from_numpy(a).to(dev) for a in (bu, bp, bn)))
    np.testing.assert_allclose(loss.cpu().numpy(), ref.losses, rtol=1e-5)
    fu, fi = om.final_embeddings()
    want, _ = O.evaluate(fu, fi, od, [20, 50], 50, mode="exact")
    got = batch_test.Test(data, model, dev, cfg)
    np.testing.assert_allclose(got["recall"], want["recall"], atol=5e-5)
    np.testing.assert_allclose(got["ndcg"], want["ndcg"], atol=5e-5)


# ---------------------------------------------------------------- foreign module through Test()
def test_Test_on_module_with_only_get_rating_for_test(dev, golden_dirs, golden_tiny):
    """SURVEY 8 b: a reference-style nn.Module that offers get_rating_for_test and nothing else evaluates to the
    reference's own Test() numbers."""
    import utility.utility_train.batch_test as batch_test
    from utility.utility_data.data_loader import Data
    g = golden_tiny
    cfg = {"dataset_path": os.path.dirname(golden_dirs["tiny"]) + "/", "dataset": "tiny", "top_K": "[10, 20]", "embedding_size": "64",
           "batch_size": "256", "test_batch_size": "50", "learn_rate": "0.001", "reg_lambda": "0.0001", "GCN_layer": "3", "sparsity_test": "0"}
    data = Data(golden_dirs["tiny"], cfg)

    class Foreign(torch.nn.Module):
        def __init__(self, fu, fi):
            super().__init__()
            self.fu, self.fi = torch.nn.Parameter(fu), torch.nn.Parameter(fi)

        def get_rating_for_test(self, user):
            with torch.no_grad():
                return torch.sigmoid(self.fu[user.long()] @ self.fi.t())

    m = Foreign(torch.from_numpy(g["lgT_fu"]).to(dev), torch.from_numpy(g["lgT_fi"]).to(dev))
    res = batch_test.Test(data, m, dev, cfg)
    np.testing.assert_allclose(res["recall"], g["lgT_test_recall"], atol=5e-5)
    np.testing.assert_allclose(res["ndcg"], g["lgT_test_ndcg"], atol=5e-5)
    np.testing.assert_allclose(res["precision"], g["lgT_test_precision"], atol=5e-5)


# ---------------------------------------------------------------- bounded peer barrier
def test_peer_barrier_times_out_with_error_code(dev):
    """A peer that never arrives surfaces as IDG_ERR_PEER_TIMEOUT + rank from idg_peers_status instead of a hung stream
    (SURVEY section 5).  One GPU plays rank 0 of a world of 2; "rank 1" is a second slab on the same device that nobody drives."""
    import ctypes as C
    from idgrec import _lib
    l = _lib.lib()
    nbytes = 1 << 20
    a, b = C.c_void_p(), C.c_void_p()
    _lib.check(l.idg_device_alloc(nbytes, C.byref(a)), "alloc")
    _lib.check(l.idg_device_alloc(nbytes, C.byref(b)), "alloc")
    bases = (C.c_void_p * 2)(a.value, b.value)
    h = C.c_void_p()
    _lib.check(l.idg_peers_create(a.value, nbytes, 0, 2, bases, C.byref(h)), "peers_create")
    _lib.check(l.idg_peers_set_timeout_ms(h, 50), "set_timeout")
    s = torch.cuda.current_stream().cuda_stream
    assert l.idg_peers_status(h, a.value, s) == 0
    _lib.check(l.idg_peers_barrier(h, a.value, s), "barrier")      # the other side never announces its epoch
    torch.cuda.synchronize()
    rc = l.idg_peers_status(h, a.value, s)
    assert rc == 100000 + 1, rc
    assert b"rank 1 did not arrive" in l.idg_last_error()
    # later barriers on the failed slab return at once
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        _lib.check(l.idg_peers_barrier(h, a.value, s), "barrier")
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 0.5
    l.idg_peers_destroy(h)
    l.idg_device_free(a); l.idg_device_free(b)


# ---------------------------------------------------------------- the reference's own main.py
REF_MAIN = "/root/reference/main.py"


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="/root/reference is only present in the build container")
def test_reference_main_py_text_runs_two_epochs(dev, golden_dirs, tmp_path):
    """Drop-in check (SURVEY 8 b): the UNMODIFIED text of the reference's main.py and Parser.py, executed from a directory
    that holds this package's modules, trains LightGCN for two epochs on the tiny dataset and logs the same losses as the
    oracle.  The two files are read from /root/reference at run time (never copied into the repo)."""
    import shutil
    work = tmp_path / "run"
    shutil.copytree(PKG, work, ignore=shutil.ignore_patterns("__pycache__"))
    for name in ("main.py", "Parser.py"):
        shutil.copy(os.path.join("/root/reference", name), work / name)
    ds = work / "dataset" / "tiny"
    ds.mkdir(parents=True)
    (work / "log").mkdir(exist_ok=True)
    shutil.copy(os.path.join(golden_dirs["tiny"], "train.txt"), ds / "train.txt")
    shutil.copy(os.path.join(golden_dirs["tiny"], "test.txt"), ds / "test.txt")
    cfg = (work / "configure" / "LightGCN.txt").read_text().splitlines()
    out = []
    for line in cfg:
        k = line.split("=")[0].strip()
        if k == "dataset": line = "dataset = tiny"
        if k == "dataset_path": line = "dataset_path = ./dataset/"
        if k == "training_epochs": line = "training_epochs = 2"
        if k == "batch_size": line = "batch_size = 256"
        if k == "test_batch_size": line = "test_batch_size = 50"
        out.append(line)
    (work / "configure" / "LightGCN.txt").write_text("\n".join(out) + "\n")
    r = subprocess.run([sys.executable, "main.py", "--model=LightGCN"], cwd=work, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Model training process completed." in r.stdout
    assert r.stdout.count("Training time:") == 2


# ---------------------------------------------------------------- row-partitioned step, degenerate world of 1
@pytest.mark.parametrize("kind,cl", [("LightGCN", 1), ("SimGCL", 1), ("XSimGCL", 1), ("XSimGCL", 2), ("XSimGCL", 3)])
@pytest.mark.parametrize("use_graph", [False, True])
def test_row_partitioned_step_world1_equals_fused_trainer(dev, kind, cl, use_graph):
    """The row-partitioned trainer (idgrec/dist.py) with a single partition must reproduce the single-GPU fused trainer bit
    for bit -- losses and tables after several steps, with the same injected noise.  Runs on one GPU, so the driver's
    1-GPU box exercises the step logic the 2-GPU tests (tests/test_gpu_dist.py) need a second device for."""
    from idgrec import datagen
    from idgrec.dist import DistFusedTrainer
    from idgrec.engine import FusedTrainer
    from idgrec.graph import Graph, build_norm_adjacency
    g = datagen.gen_graph("small")
    U, I = g.num_users, g.num_items
    csr = build_norm_adjacency(g.train_user, g.train_item, U, I, device=dev)
    G = Graph(csr)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    table = (torch.rand(U + I, 64, generator=gen, device=dev) - 0.5) * 0.2
    kw = dict(ssl_lambda=0.3, temperature=0.2, eps=0.1, cl_layer=cl) if kind != "LightGCN" else {}
    ref = FusedTrainer(kind, G, table.clone(), U, 3, 1e-4, 1e-3, max_batch=512, use_cuda_graph=False, **kw)
    ft = DistFusedTrainer(kind, csr, table.clone(), U, 3, 1e-4, 1e-3, 0, 1, max_batch=512, use_cuda_graph=use_graph, full_graph=G, **kw)
    n_views = {"LightGCN": 0, "SimGCL": 2, "XSimGCL": 1}[kind]
    rng = np.random.default_rng(8)
    nz = [torch.empty(3, U + I, 64, device=dev) for _ in range(n_views)]   # stable buffers: a captured step replays the same pointers
    ref.injected_noise = ft.injected_noise = nz if n_views else None
    for step in range(3):
        e = rng.integers(0, len(g.train_user), 512)
        b = tuple(torch.from_numpy(a).to(dev) for a in (g.train_user[e], g.train_item[e], rng.integers(0, I, 512)))
        for t in nz:
            t.copy_(torch.rand(3, U + I, 64, generator=gen, device=dev))
        lr, ld = ref.step(*b).clone(), ft.step(*b).clone()
        assert torch.equal(lr, ld), (step, lr, ld)
    assert torch.equal(ref.E0, ft.E0)
    fr = G.propagate_fwd(ref.E0, 3, kind == "LightGCN")
    assert torch.equal(fr, ft.final_embeddings())


# ---------------------------------------------------------------- tensor-core candidate pass: per-item upper bounds
@pytest.mark.parametrize("d", [64, 256])
def test_tc_candidate_upper_bounds_carry_the_per_item_margin(dev, d):
    """The tcgen05 candidate pass keeps item i for user u iff w_ui >= L_u, where w_ui = s~_ui + c|u||i| comes out of the
    accumulator (ninth k-step: margin operand).  Direct check of that operand on heavy-tailed item norms: the dumped
    accumulators must (1) dominate the exact fp64 scores everywhere and (2) exceed the product of the tf32-rounded operands by
    exactly the per-item margin (and that product must itself be within the bound of the exact score).  Without this the exact-rank tests could not tell a silently missing margin from a working one."""
    import ctypes as C
    from idgrec import _lib
    l = _lib.lib()
    U, I = 128, 640           # d = 256 (NGCF's concatenated layers): four k-chunks per item tile, margin with the last
    gen = torch.Generator().manual_seed(17)
    Fu = (torch.randn(U, d, generator=gen) * 0.5).numpy()
    Fi = (torch.randn(I, d, generator=gen) * 0.3).numpy()
    Fi[::7] *= 8.0            # heavy-tailed norms, like a trained table (largest norm ~8x the median)
    Fi[3] = 0.0               # a zero row: zero margin, zero score
    users = np.arange(U, dtype=np.int64)[::-1].copy()
    mp = torch.zeros(U + 1, dtype=torch.int32, device=dev)
    mi = torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.full((128, 128), float("nan"), dtype=torch.float32, device=dev)
    ws = torch.empty(int(l.idg_eval_workspace_bytes(U, I, d, 20)), dtype=torch.uint8, device=dev)
    fu_d, fi_d, us_d = torch.from_numpy(Fu).to(dev), torch.from_numpy(Fi).to(dev), torch.from_numpy(users).to(dev)
    _lib.check(l.idg_eval_tc_bounds(fu_d.data_ptr(), fi_d.data_ptr(), U, I, d, mp.data_ptr(), mi.data_ptr(), us_d.data_ptr(), U, out.data_ptr(),
                                    ws.data_ptr(), torch.cuda.current_stream().cuda_stream), "idg_eval_tc_bounds")
    w = out.cpu().numpy().astype(np.float64)
    assert np.isfinite(w).all()
    exact = O.scores_fp64_sequential(Fu[users], Fi[:128])
    assert (w >= exact).all(), "an accumulator is below the exact score: the upper bound does not hold"
    # cvt.rna.tf32: round to nearest (ties away) on the 13 dropped mantissa bits -- what both operands go through before the MMA
    tf = lambda a: ((a.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32).astype(np.float64)
    approx = tf(Fu[users].copy()) @ tf(Fi[:128].copy()).T
    nu_, ni_ = np.linalg.norm(Fu[users].astype(np.float64), axis=1), np.linalg.norm(Fi[:128].astype(np.float64), axis=1)
    margin = 1.03 * (2.0 ** -10 + 2.4e-7 + (3.9e-6 if d == 64 else 1.54e-5)) * nu_[:, None] * ni_[None, :]
    assert (np.abs(approx - exact) <= margin / 1.03 + 1e-30).all(), "the rounded-operand product is further from the exact score than the bound"
    got = w - approx
    assert np.all(np.abs(got - margin) <= 0.01 * margin + 8e-6 * nu_[:, None] * ni_[None, :] + 1e-30), float(np.abs(got - margin).max())
    assert np.all(got[:, 3] == 0.0)


@pytest.mark.parametrize("scale", [0.05, 0.3])
def test_eval_topk_exact_rank_width_256_on_tensor_cores(dev, scale):
    """NGCF ranks on the concatenation of its layers (models/NGCF.py:108,132-138: 64 + 64*3 = 256 columns).  That width goes
    through the tcgen05 candidate pass too (user tile resident, item tiles as four 64-wide k-chunks): ids bit-exact against the
    fp64 exact-rank oracle, several item tiles, a ragged last tile, duplicated rows (ties by id), heavy-tailed norms."""
    from idgrec import ops
    U, I, d, K = 300, 128 * 9 + 37, 256, 20
    net = _rand_net(U, I, 9000, 256)
    gen = torch.Generator().manual_seed(256)
    Fu = (torch.randn(U, d, generator=gen) * scale).numpy()
    Fi = (torch.randn(I, d, generator=gen) * scale).numpy()
    Fi[::11] *= 4.0
    Fi[200:215] = Fi[200]
    users = np.arange(U, dtype=np.int64)[::-1].copy()
    ref_ids, ref_sc = O.topk_exact(Fu, Fi, users, net.indptr, net.indices, K)
    mp, mi = _mask(net, dev)
    ids, sc = ops.eval_topk(torch.from_numpy(Fu).to(dev), torch.from_numpy(Fi).to(dev), torch.from_numpy(users).to(dev), mp, mi, K, want_scores=True)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_ids)
    np.testing.assert_allclose(sc.cpu().numpy(), ref_sc, rtol=1e-6)


# ---------------------------------------------------------------- NGCF: one-launch dropout draws
def test_ngcf_keep_masks_are_bernoulli_and_step_dependent(dev):
    """idg_ngcf_keep_masks replaces three torch launches per mask (NGCF.py:99-100's nn.Dropout draw): per-layer keep rates match
    1 - p, the draw depends on (seed, step) only -- same inputs, same mask; next step, fresh mask -- and the entries are 0/1."""
    import ctypes as C
    from idgrec import _lib
    l = _lib.lib()
    per, K = 1 << 20, 3
    keep = torch.empty(K, per, device=dev)
    probs = (C.c_float * K)(0.9, 0.5, 1.0)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(l.idg_ngcf_keep_masks(keep.data_ptr(), per, K, probs, 1234, step.data_ptr(), s), "keep_masks")
    a = keep.clone()
    assert set(torch.unique(a).tolist()) <= {0.0, 1.0}
    rates = a.mean(dim=1).cpu().numpy()
    np.testing.assert_allclose(rates, [0.9, 0.5, 1.0], atol=3e-3)
    _lib.check(l.idg_ngcf_keep_masks(keep.data_ptr(), per, K, probs, 1234, step.data_ptr(), s), "keep_masks")
    assert torch.equal(a, keep)
    step += 1
    _lib.check(l.idg_ngcf_keep_masks(keep.data_ptr(), per, K, probs, 1234, step.data_ptr(), s), "keep_masks")
    assert not torch.equal(a[0], keep[0]) and abs(float((a[0] * keep[0]).mean()) - 0.81) < 5e-3      # independent draws


@pytest.mark.parametrize("kind,use_graph", [("LightGCN", True), ("LightGCN", False), ("SimGCL", True), ("XSimGCL", True)])
def test_row_partitioned_step_world1_chunked_exchange(dev, monkeypatch, kind, use_graph):
    """The chunked exchange (local rows computed in nnz-balanced blocks on per-block propagation handles, each finished block pushed
    from a second stream -- the default at 8 ranks on large graphs) with a single partition: same bits as the single-GPU fused
    trainer.  The pushes are no-ops in a world of one; the block schedule, the per-block handles and the fork / join are real."""
    from idgrec import datagen
    from idgrec.dist import DistFusedTrainer
    from idgrec.engine import FusedTrainer
    from idgrec.graph import Graph, build_norm_adjacency
    monkeypatch.setenv("IDG_DIST_EXCHANGE", "chunked")
    monkeypatch.setenv("IDG_DIST_CHUNKS", "5")
    g = datagen.gen_graph("small")
    U, I = g.num_users, g.num_items
    csr = build_norm_adjacency(g.train_user, g.train_item, U, I, device=dev)
    G = Graph(csr)
    gen = torch.Generator(device=dev)
    gen.manual_seed(15)
    table = (torch.rand(U + I, 64, generator=gen, device=dev) - 0.5) * 0.2
    kw = dict(ssl_lambda=0.3, temperature=0.2, eps=0.1, cl_layer=2) if kind != "LightGCN" else {}
    ref = FusedTrainer(kind, G, table.clone(), U, 3, 1e-4, 1e-3, max_batch=512, use_cuda_graph=False, **kw)
    ft = DistFusedTrainer(kind, csr, table.clone(), U, 3, 1e-4, 1e-3, 0, 1, max_batch=512, use_cuda_graph=use_graph, full_graph=G, **kw)
    assert ft.chunked and len(ft.chunks) >= 4
    n_views = {"LightGCN": 0, "SimGCL": 2, "XSimGCL": 1}[kind]
    rng = np.random.default_rng(18)
    nz = [torch.empty(3, U + I, 64, device=dev) for _ in range(n_views)]
    ref.injected_noise = ft.injected_noise = nz if n_views else None
    for step in range(3):
        e = rng.integers(0, len(g.train_user), 512)
        b = tuple(torch.from_numpy(a).to(dev) for a in (g.train_user[e], g.train_item[e], rng.integers(0, I, 512)))
        for t in nz:
            t.copy_(torch.rand(3, U + I, 64, generator=gen, device=dev))
        lr, ld = ref.step(*b).clone(), ft.step(*b).clone()
        assert torch.equal(lr, ld), (step, lr, ld)
    assert torch.equal(ref.E0, ft.E0)


@pytest.mark.parametrize("use_graph", [False, True])
def test_row_partitioned_step_world1_with_neighbourhood_restriction(dev, use_graph):
    """Same bit-identity with the batch-neighbourhood (closure) restriction forced on: masked layer K-1, row-masked sparse first
    backward product, closure-column second product, closure-based re-zeroing -- against the DENSE single-GPU fused trainer."""
    from idgrec import datagen
    from idgrec.dist import DistFusedTrainer
    from idgrec.engine import FusedTrainer
    from idgrec.graph import Graph, build_norm_adjacency
    g = datagen.gen_graph("small")
    U, I = g.num_users, g.num_items
    csr = build_norm_adjacency(g.train_user, g.train_item, U, I, device=dev)
    G = Graph(csr)
    gen = torch.Generator(device=dev)
    gen.manual_seed(6)
    table = (torch.rand(U + I, 64, generator=gen, device=dev) - 0.5) * 0.2
    ref = FusedTrainer("LightGCN", G, table.clone(), U, 3, 1e-4, 1e-3, max_batch=64, use_cuda_graph=False, restrict_rows=False)
    ref2 = FusedTrainer("LightGCN", Graph(csr), table.clone(), U, 3, 1e-4, 1e-3, max_batch=64, use_cuda_graph=use_graph, closure_restrict=True)
    ft = DistFusedTrainer("LightGCN", csr, table.clone(), U, 3, 1e-4, 1e-3, 0, 1, max_batch=64, use_cuda_graph=use_graph, closure_restrict=True)
    assert ft.use_closure and ref2.use_closure
    rng = np.random.default_rng(9)
    for step in range(4):
        e = rng.integers(0, len(g.train_user), 64)          # a small batch: its neighbourhood is a small part of the graph
        b = tuple(torch.from_numpy(a).to(dev) for a in (g.train_user[e], g.train_item[e], rng.integers(0, I, 64)))
        l0, l1, l2 = ref.step(*b).clone(), ref2.step(*b).clone(), ft.step(*b).clone()
        assert torch.equal(l0, l1) and torch.equal(l0, l2), (step, l0, l1, l2)
    assert torch.equal(ref.E0, ref2.E0) and torch.equal(ref.E0, ft.E0)


# ---------------------------------------------------------------- contrastive steps against the fp64 evaluation of the reference formulas
@pytest.mark.parametrize("kind,cl", [("SimGCL", 1), ("XSimGCL", 1), ("XSimGCL", 3)])
def test_contrastive_step_losses_and_gradient_vs_fp64(dev, kind, cl):
    """The golden comparisons of the contrastive models state 1e-4 on gradients because the goldens are the reference's own fp32
    CPU autograd, which is itself ~1e-4 from the exact value.  Against the SAME formulas (oracle/ref_oracle.py: propagate,
    bpr_loss, reg_loss, infonce_loss) evaluated in float64 the fused CUDA step is held to 1e-5 on the three losses and 1e-5 of
    the largest entry on the ego-table gradient (measured: 0.7e-6 .. 2.6e-6)."""
    import scipy.sparse as sp
    from idgrec import datagen
    from idgrec.engine import FusedTrainer
    from idgrec.graph import Graph, build_norm_adjacency
    g = datagen.gen_graph("small")
    U, I, K, B = g.num_users, g.num_items, 3, 512
    N = U + I
    csr = build_norm_adjacency(g.train_user, g.train_item, U, I, device=dev)
    n_views = 2 if kind == "SimGCL" else 1
    eps = 0.1
    net = sp.csr_matrix((np.ones(len(g.train_user)), (g.train_user, g.train_item)), shape=(U, I))
    net.sort_indices()
    ip, ix, dt, _ = O.norm_adjacency(net)
    A = O.csr_to_torch_coo(ip, ix, dt, N).double()
    # x += sign(x) * ... is discontinuous at x = 0: an entry whose fp32 and fp64 values straddle zero flips a whole noise component
    # (observed: one flip in 1.9 M entries moved the gradient by 4e-3).  Draw inputs until no pre-noise entry of a perturbed
    # layer is within 1e-8 of zero (fp32 rounding of these sums is ~1e-9), so the comparison measures arithmetic, not the jump.
    for seed in range(12, 40):
        gen = torch.Generator().manual_seed(seed)
        table = (torch.rand(N, 64, generator=gen) - 0.5) * 0.2
        noise = [torch.rand(K, N, 64, generator=gen) for _ in range(n_views)]
        closest = float("inf")
        for v in noise:
            x = table.double()
            for k in range(K):
                x = torch.sparse.mm(A, x)
                closest = min(closest, float(x.abs()[x != 0].min()))      # rows without neighbours are exactly 0 on both sides
                x = x + torch.sign(x) * torch.nn.functional.normalize(v[k].double(), dim=-1) * eps
        if closest > 1e-8:
            break
    else:
        pytest.skip("no draw without a near-zero pre-noise entry")
    rng = np.random.default_rng(4)
    e = rng.integers(0, len(g.train_user), B)
    bu, bp, bn = g.train_user[e], g.train_item[e], rng.integers(0, I, B)
    tau, lam, reg = 0.2, 0.3, 1e-4
    # ---- float64 evaluation of the reference formulas
    X0 = table.double().requires_grad_(True)
    u, p, n = (torch.as_tensor(a, dtype=torch.long) for a in (bu, bp, bn))
    nz = [[t.double() for t in v] for v in noise]
    if kind == "SimGCL":
        F = O.propagate(A, X0, K, False)
        V1, V2 = O.propagate(A, X0, K, False, nz[0], eps), O.propagate(A, X0, K, False, nz[1], eps)
    else:
        F, V1 = O.propagate(A, X0, K, False, nz[0], eps, cl)
        V2 = F
    bpr = O.bpr_loss(F[u], F[U + p], F[U + n])
    regl = reg * O.reg_loss(X0[u], X0[U + p], X0[U + n])
    ui, ii = torch.unique(u), torch.unique(p) + U
    ssl = lam * (O.infonce_loss(V1[ui], V2[ui], tau) + O.infonce_loss(V1[ii], V2[ii], tau))
    (bpr + regl + ssl).backward()
    want_loss, want_grad = np.array([bpr.item(), regl.item(), ssl.item()]), X0.grad.numpy()
    # ---- fused CUDA step
    ft = FusedTrainer(kind, Graph(csr), table.to(dev), U, K, reg, 1e-3, max_batch=B, use_cuda_graph=False, ssl_lambda=lam, temperature=tau, eps=eps, cl_layer=cl)
    ft.injected_noise = [v.to(dev) for v in noise]
    got = ft.step(*(torch.from_numpy(a).to(dev) for a in (bu, bp, bn)), apply_adam=False).cpu().numpy()
    np.testing.assert_allclose(got, want_loss, rtol=1e-5)
    err = np.abs(ft.gE0.cpu().numpy().astype(np.float64) - want_grad).max() / np.abs(want_grad).max()
    print("%s cl=%d: gradient max error / max entry = %.2e" % (kind, cl, err))
    assert err <= 1e-5, err


# ---------------------------------------------------------------- NGCF dense layer on tcgen05: forward and backward against fp64
@pytest.mark.parametrize("N,with_keep,with_ext", [(1000, True, True), (777, True, False), (130, False, True), (19001, True, True)])
def test_ngcf_dense_layer_tensor_core_kernels_vs_fp64(dev, N, with_keep, with_ext):
    """models/NGCF.py:87-106 for one layer, forward (persistent tcgen05 kernel, csrc/ngcf_tc.cu) and its autograd (csrc/ngcf_bwd_tc.cu:
    dZ^T = Wcat . dS^T with Wcat in TMEM, dWcat = Z^T . dS from MN-major operands) against an fp64 restatement: every output within
    1e-5 of the largest reference entry.  Ragged row counts (last tile partial, fewer tiles than SMs, several tiles per CTA), no
    dropout mask, no incoming gradient from a next layer; the forward must not touch the other blocks of the [N,256] concat."""
    from idgrec import _lib
    from idgrec._lib import check, ptr, cur_stream
    l = _lib.lib()
    g = torch.Generator(device=dev).manual_seed(N)
    rn = lambda *s: torch.randn(*s, generator=g, device=dev)
    E, side = rn(N, 64) * 0.3, rn(N, 64) * 0.3
    Wg, Wb, bg, bb = rn(64, 64) * 0.2, rn(64, 64) * 0.2, rn(64) * 0.1, rn(64) * 0.1
    keep = (torch.rand(N, 64, generator=g, device=dev) < 0.9).float() if with_keep else None
    p = 0.1 if with_keep else 0.0
    D = torch.empty(N, 64, device=dev)
    out = torch.zeros(N, 256, device=dev)
    kp = ptr(keep) if with_keep else None
    check(l.idg_ngcf_dense_fwd(ptr(E), ptr(side), ptr(Wg), ptr(bg), ptr(Wb), ptr(bb), kp, p, N, None, ptr(D), out[:, 128:].data_ptr(), 256, cur_stream()), "fwd")
    f = lambda t: t.double()
    kd = f(keep) if with_keep else 1.0
    Sd = f(side) @ f(Wg) + f(bg) + (f(E) * f(side)) @ f(Wb) + f(bb)
    Dd = torch.where(Sd > 0, Sd, 0.2 * Sd) * kd / (1 - p)
    nrm = Dd.norm(dim=1, keepdim=True).clamp_min(1e-12)
    Od = Dd / nrm
    rel = lambda got, ref: float((f(got) - ref).abs().max() / ref.abs().max())
    assert rel(D, Dd) < 1e-5 and rel(out[:, 128:192], Od) < 1e-5
    assert float(out[:, :128].abs().max()) == 0.0 and float(out[:, 192:].abs().max()) == 0.0
    # backward, from the kernel's own D (what the fused step passes)
    dO_full, dDx = rn(N, 256), (rn(N, 64) * 0.5 if with_ext else None)
    dO = dO_full[:, 128:192]
    dside, dEd = torch.empty(N, 64, device=dev), torch.empty(N, 64, device=dev)
    dWg, dWb, db = torch.empty(64, 64, device=dev), torch.empty(64, 64, device=dev), torch.empty(64, device=dev)
    ws = torch.empty(int(l.idg_ngcf_workspace_bytes()), dtype=torch.uint8, device=dev)
    check(l.idg_ngcf_dense_bwd(ptr(E), ptr(side), ptr(Wg), ptr(Wb), kp, p, None, ptr(D), dO.data_ptr(), 256, ptr(dDx) if with_ext else None, N, ptr(dside),
                               ptr(dEd), ptr(dWg), ptr(dWb), ptr(db), ptr(ws), cur_stream()), "bwd")
    D64 = f(D)
    n64 = D64.norm(dim=1, keepdim=True).clamp_min(1e-12)
    O64 = D64 / n64
    dD = (f(dDx) if with_ext else 0.0) + (f(dO) - O64 * (O64 * f(dO)).sum(1, keepdim=True)) / n64
    dS = dD * kd / (1 - p) * torch.where(D64 > 0, 1.0, 0.2)
    Wcat, Z = torch.cat([f(Wg), f(Wb)], 0), torch.cat([f(side), f(E) * f(side)], 1)
    dZ = dS @ Wcat.T
    dW = Z.T @ dS
    for name, got, ref in (("dside", dside, dZ[:, :64] + dZ[:, 64:] * f(E)), ("dE_direct", dEd, dZ[:, 64:] * f(side)), ("dWg", dWg, dW[:64]),
                           ("dWb", dWb, dW[64:]), ("db", db, dS.sum(0))):
        assert rel(got, ref) < 1e-5, (name, rel(got, ref))


def test_ngcf_bit_packed_dropout_masks_match_the_float_masks(dev):
    """idg_ngcf_keep_bits packs the SAME Philox draws as idg_ngcf_keep_masks (64 bits per row), and the dense kernels give the same
    bits whether the mask arrives as floats or as bits (forward D / O, backward dside / dE_direct / dW / db)."""
    import ctypes as C
    from idgrec import _lib
    from idgrec._lib import check, ptr, cur_stream
    l = _lib.lib()
    N, K = 1500, 2
    probs = (C.c_float * K)(0.9, 0.6)
    step = torch.full((1,), 3, dtype=torch.int32, device=dev)
    keep = torch.empty(K, N, 64, device=dev)
    bits = torch.zeros(K, N, 2, dtype=torch.int32, device=dev)
    s = cur_stream()
    check(l.idg_ngcf_keep_masks(ptr(keep), N * 64, K, probs, 77, ptr(step), s), "keep_masks")
    check(l.idg_ngcf_keep_bits(ptr(bits), N, K, probs, 77, ptr(step), s), "keep_bits")
    w = bits.cpu().numpy().astype(np.uint32)                                    # [K, N, 2]
    unpacked = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(K, N, 64).astype(np.float32)
    np.testing.assert_array_equal(unpacked, keep.cpu().numpy())
    g = torch.Generator(device=dev).manual_seed(1)
    rn = lambda *sh: torch.randn(*sh, generator=g, device=dev)
    E, side, Wg, Wb, bg, bb = rn(N, 64) * 0.3, rn(N, 64) * 0.3, rn(64, 64) * 0.2, rn(64, 64) * 0.2, rn(64) * 0.1, rn(64) * 0.1
    dO, dDx = rn(N, 64), rn(N, 64) * 0.5
    ws = torch.empty(int(l.idg_ngcf_workspace_bytes()), dtype=torch.uint8, device=dev)
    res = []
    for use_bits in (False, True):
        D, out = torch.empty(N, 64, device=dev), torch.empty(N, 64, device=dev)
        outs = [torch.empty(N, 64, device=dev), torch.empty(N, 64, device=dev), torch.empty(64, 64, device=dev), torch.empty(64, 64, device=dev),
                torch.empty(64, device=dev)]
        if use_bits:
            check(l.idg_ngcf_dense_fwd_bits(ptr(E), ptr(side), ptr(Wg), ptr(bg), ptr(Wb), ptr(bb), ptr(bits[1]), 0.4, N, None, ptr(D), ptr(out), 64, s), "fwd_bits")
            check(l.idg_ngcf_dense_bwd_bits(ptr(E), ptr(side), ptr(Wg), ptr(Wb), ptr(bits[1]), 0.4, ptr(D), ptr(dO), 64, ptr(dDx), N, *[ptr(o) for o in outs],
                                            ptr(ws), s), "bwd_bits")
        else:
            check(l.idg_ngcf_dense_fwd(ptr(E), ptr(side), ptr(Wg), ptr(bg), ptr(Wb), ptr(bb), ptr(keep[1]), 0.4, N, None, ptr(D), ptr(out), 64, s), "fwd")
            check(l.idg_ngcf_dense_bwd(ptr(E), ptr(side), ptr(Wg), ptr(Wb), ptr(keep[1]), 0.4, None, ptr(D), ptr(dO), 64, ptr(dDx), N, *[ptr(o) for o in outs],
                                       ptr(ws), s), "bwd")
        res.append([D, out] + outs)
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_step_graph_fetches_epoch_slices_itself(dev):
    """trainer.py:40-47 walks the epoch's sample arrays in slices.  When step() is handed such slices in order, the captured step
    fetches its batch on the device (idg_batch_fetch as the first graph node); any other batch goes through the copy path.  Same
    tables and losses either way, across an epoch boundary, a ragged last batch and an out-of-order call in the middle."""
    from idgrec import datagen
    from idgrec.engine import FusedTrainer
    from idgrec.graph import Graph, build_norm_adjacency
    g = datagen.gen_graph("small")
    U, I = g.num_users, g.num_items
    csr = build_norm_adjacency(g.train_user, g.train_item, U, I, device=dev)
    gen = torch.Generator(device=dev).manual_seed(21)
    table = (torch.rand(U + I, 64, generator=gen, device=dev) - 0.5) * 0.2
    B, E = 256, 256 * 5 + 77
    rng = np.random.default_rng(4)
    a = FusedTrainer("LightGCN", Graph(csr), table.clone(), U, 3, 1e-4, 1e-3, max_batch=B, use_cuda_graph=True)
    b = FusedTrainer("LightGCN", Graph(csr), table.clone(), U, 3, 1e-4, 1e-3, max_batch=B, use_cuda_graph=True)
    from idgrec.dist import DistFusedTrainer
    c = DistFusedTrainer("LightGCN", csr, table.clone(), U, 3, 1e-4, 1e-3, 0, 1, max_batch=B, use_cuda_graph=True)     # row-partitioned trainer, one partition
    for epoch in range(2):
        e = rng.integers(0, len(g.train_user), E)
        block = torch.from_numpy(np.stack([g.train_user[e], g.train_item[e], rng.integers(0, I, E)])).to(dev)      # [3, E], like trainer.py's
        users, pos, neg = block[0], block[1], block[2]
        for s in range(0, E, B):
            sl = (users[s:s + B], pos[s:s + B], neg[s:s + B])
            la = a.step(*sl).clone()
            lb = b.step(*[t.clone() for t in sl]).clone()          # clones are not views: always the copy path
            lc = c.step(*sl).clone()
            assert torch.equal(la, lb) and torch.equal(la, lc), (epoch, s)
            if epoch == 1 and s == 2 * B:                          # an extra, out-of-order batch: a falls back to the copy path from here on
                la, lb = a.step(users[:B], pos[:B], neg[:B]).clone(), b.step(users[:B].clone(), pos[:B].clone(), neg[:B].clone()).clone()
                lc = c.step(users[:B], pos[:B], neg[:B]).clone()
                assert torch.equal(la, lb) and torch.equal(la, lc)
    assert any(isinstance(k, tuple) for k in a._graphs) and not any(isinstance(k, tuple) for k in b._graphs)
    assert ("f", 77, B) in a._graphs
    assert any(isinstance(k, tuple) for k in c._graphs)
    assert torch.equal(a.E0, b.E0) and torch.equal(a.E0, c.E0)


# ---------------------------------------------------------------- class-split schedule (tables beyond the L2)
@pytest.mark.parametrize("closure", [False, True])
def test_class_split_schedule_is_bit_identical(dev, monkeypatch, closure):
    """idg_graph_create orders the work items "all user rows, then all item rows" for gather tables that do not fit the L2
    (csrc/spmm.cu:classify_rows; forced here on the small graph).  The order is scheduling only: one propagation layer, the
    row-restricted / sparse-input / Adam-fused layers of the fused step (through row_items and the work lists) and the heavy rows'
    chunk-order partial sums give the same bits as the single schedule.  With self loops no row is purely upper or lower and the
    single schedule stays."""
    from idgrec import datagen
    from idgrec.engine import FusedTrainer
    from idgrec.graph import Graph, build_norm_adjacency
    g = datagen.gen_graph("small")
    U, I = g.num_users, g.num_items
    csr = build_norm_adjacency(g.train_user, g.train_item, U, I, device=dev)
    monkeypatch.setenv("IDG_SPMM_CLASS_SPLIT", "0")
    G0 = Graph(csr)
    monkeypatch.setenv("IDG_SPMM_CLASS_SPLIT", "1")
    G1 = Graph(csr)
    csr_self = build_norm_adjacency(g.train_user, g.train_item, U, I, add_self=True, device=dev)
    Gs1 = Graph(csr_self)
    monkeypatch.setenv("IDG_SPMM_CLASS_SPLIT", "0")
    Gs0 = Graph(csr_self)
    monkeypatch.delenv("IDG_SPMM_CLASS_SPLIT")
    from idgrec import _lib
    L = _lib.lib()
    assert [L.idg_graph_classes(h._h) for h in (G0, G1, Gs0, Gs1)] == [1, 2, 1, 1]
    deg = np.diff(csr.indptr.cpu().numpy())
    assert deg.max() > 256, "the small graph must hold heavy rows for this test to cover the partial slots"
    gen = torch.Generator(device=dev).manual_seed(33)
    X = (torch.rand(U + I, 64, generator=gen, device=dev) - 0.5) * 0.2
    for ga, gb in ((G0, G1), (Gs0, Gs1)):
        Ya, Yb = torch.empty_like(X), torch.empty_like(X)
        ga.spmm_layer(X, Y=Ya)
        gb.spmm_layer(X, Y=Yb)
        assert torch.equal(Ya, Yb) and float(Ya.abs().sum()) > 0
    B = 64 if closure else 256
    kw = dict(closure_restrict=True) if closure else {}
    monkeypatch.setenv("IDG_SPMM_ADAM8", "0")
    a = FusedTrainer("LightGCN", G0, X.clone(), U, 3, 1e-4, 1e-3, max_batch=B, use_cuda_graph=False, **kw)
    b = FusedTrainer("LightGCN", G1, X.clone(), U, 3, 1e-4, 1e-3, max_batch=B, use_cuda_graph=True, **kw)
    rng = np.random.default_rng(5)
    for step in range(3):
        e = rng.integers(0, len(g.train_user), B)
        batch = tuple(torch.from_numpy(t).to(dev) for t in (g.train_user[e], g.train_item[e], rng.integers(0, I, B)))
        la = a.step(*batch).clone()
        monkeypatch.setenv("IDG_SPMM_ADAM8", "1")      # b also runs the 64-warp build of the Adam-fused layer (the XL-table choice)
        lb = b.step(*batch).clone()
        monkeypatch.setenv("IDG_SPMM_ADAM8", "0")
        assert torch.equal(la, lb), (step, la, lb)
    assert torch.equal(a.E0, b.E0)

