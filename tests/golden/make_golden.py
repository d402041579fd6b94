"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so the
pins are outputs of its own modules -- Data, data_graph, tools, losses, metrics,
models.{LightGCN,SimGCL,XSimGCL,NGCF,MFBPR}, batch_test.Test -- on small
synthetic datasets written in its file format.  Nothing from the reference is
copied: it is imported from where it lies and only its *outputs* are stored.
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
PKG = os.path.join(REPO, "id-grec_b200")
sys.path.insert(0, PKG)
from idgrec import datagen  # noqa: E402  (only the dataset writer; no product compute)

# the product package mirrors the reference's module names (utility.*, models.*, Parser) and the reference's
# directories are namespace packages (no __init__.py), so a regular package of the same name anywhere on sys.path
# would shadow them: take the product directory off the path again before importing the reference
sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (PKG, REPO)]
for m in [k for k in sys.modules if k.split(".")[0] in ("utility", "models", "Parser")]:
    del sys.modules[m]
sys.path.insert(0, REF)
import utility.utility_data.data_loader as ref_loader  # noqa: E402
import utility.utility_data.data_graph as ref_graph  # noqa: E402
import utility.utility_function.tools as ref_tools  # noqa: E402
import utility.utility_function.losses as ref_losses  # noqa: E402
import utility.utility_function.metrics as ref_metrics  # noqa: E402
import utility.utility_train.batch_test as ref_test  # noqa: E402

assert ref_loader.__file__.startswith(REF), ref_loader.__file__


def ragged(list_of_arrays):
    ptr = np.cumsum([0] + [len(a) for a in list_of_arrays]).astype(np.int64)
    flat = np.concatenate([np.asarray(a, dtype=np.int64) for a in list_of_arrays]) if ptr[-1] else np.zeros(0, np.int64)
    return ptr, flat


def canon_csr(m):
    m = m.tocsr().copy()
    m.sort_indices()
    return m.indptr, m.indices, m.data


def base_config(name, root, **kw):
    c = {"dataset_path": root + "/", "dataset": name, "top_K": "[10, 20]", "training_epochs": "2",
         "early_stopping": "10", "interval": "1", "embedding_size": "64", "batch_size": "256",
         "test_batch_size": "50", "learn_rate": "0.001", "reg_lambda": "0.0001", "GCN_layer": "3",
         "sparsity_test": "0"}
    c.update({k: str(v) for k, v in kw.items()})
    return c


class RecordingRand:
    """Captures every torch.rand_like draw of the reference's perturbation."""

    def __init__(self):
        self.draws = []
        self.orig = torch.rand_like

    def __call__(self, x, *a, **k):
        r = self.orig(x, *a, **k)
        self.draws.append(r.clone())
        return r


class RecordingDropout(torch.nn.Module):
    masks = []

    def __init__(self, p):
        super().__init__()
        self.p = p

    def forward(self, x):
        keep = (torch.rand_like(x) >= self.p).to(x.dtype)
        RecordingDropout.masks.append(keep.clone())
        return x * keep / (1.0 - self.p)


def run_dataset(name, root, out):
    cfg = base_config(name, root)
    ref_tools.set_seed(2024)
    data = ref_loader.Data(root + "/" + name, cfg)
    out["num_users"], out["num_items"] = data.num_users, data.num_items
    out["num_train"], out["num_test"] = data.num_train, data.num_test
    out["stats"] = np.array(data.get_statistics())
    out["train_user"], out["train_item"] = data.train_user, data.train_item
    out["allpos_ptr"], out["allpos_flat"] = ragged(data.all_positive)
    tu = list(data.test_dict.keys())
    out["test_users"] = np.array(tu, dtype=np.int64)
    out["test_ptr"], out["test_flat"] = ragged([data.test_dict[u] for u in tu])
    out["net_indptr"], out["net_indices"], out["net_data"] = canon_csr(data.user_item_net)

    # a2: both adjacency variants straight from the reference builders
    A = ref_graph.sparse_adjacency_matrix(data)
    out["A_dtype"] = np.array(str(A.dtype))
    out["A_sorted"] = bool(A.has_sorted_indices)
    out["A_indptr"], out["A_indices"], out["A_data"] = canon_csr(A)
    As = ref_graph.sparse_adjacency_matrix_with_self(data)
    out["As_dtype"] = np.array(str(As.dtype))
    t = ref_tools.convert_sp_mat_to_sp_tensor(As).coalesce()  # tools.py:95-109 + NGCF.py:51
    out["As_coo_index"], out["As_coo_value"] = t.indices().numpy(), t.values().numpy()
    t = ref_tools.convert_sp_mat_to_sp_tensor(A).coalesce()
    out["A_coo_index"], out["A_coo_value"] = t.indices().numpy(), t.values().numpy()

    # a4/a5: two epochs of sampler + shuffle on one RNG stream
    ref_tools.set_seed(2024)
    for ep in range(2):
        s = data.sample_data_to_train_all()
        out["sample_ep%d" % ep] = s
        perm = ref_tools.shuffle(np.arange(len(s)))
        out["perm_ep%d" % ep] = perm
    out["rng_after_pos"] = np.random.get_state()[2]
    out["rng_after_key"] = np.random.get_state()[1]
    return data, cfg


def run_models(name, data, cfg, out):
    dev = torch.device("cpu")
    B = int(cfg["batch_size"])
    samples = out["sample_ep0"][out["perm_ep0"]]
    bu, bp, bn = (torch.from_numpy(samples[:B, j].copy()).long() for j in range(3))
    out["batch"] = samples[:B]
    bu2, bp2, bn2 = (torch.from_numpy(samples[B:2 * B, j].copy()).long() for j in range(3))

    # ---- LightGCN: init, aggregate, 2 train steps, eval
    from models.LightGCN import LightGCN
    ref_tools.set_seed(2024)
    m = LightGCN(cfg, data, dev)
    out["lg_user_w0"], out["lg_item_w0"] = m.user_embedding.weight.detach().numpy().copy(), m.item_embedding.weight.detach().numpy().copy()
    fu, fi = m.aggregate()
    out["lg_fu0"], out["lg_fi0"] = fu.detach().numpy().copy(), fi.detach().numpy().copy()
    opt = torch.optim.Adam(m.parameters(), lr=float(cfg["learn_rate"]))
    for st, (a, b, c) in enumerate([(bu, bp, bn), (bu2, bp2, bn2)]):
        ll = m(a, b, c)
        tot = sum(ll)
        opt.zero_grad()
        tot.backward()
        out["lg_loss_s%d" % st] = np.array([l.item() for l in ll])
        out["lg_gu_s%d" % st] = m.user_embedding.weight.grad.numpy().copy()
        out["lg_gi_s%d" % st] = m.item_embedding.weight.grad.numpy().copy()
        opt.step()
        out["lg_user_w_s%d" % st] = m.user_embedding.weight.detach().numpy().copy()
        out["lg_item_w_s%d" % st] = m.item_embedding.weight.detach().numpy().copy()
    res = ref_test.Test(data, m, dev, cfg)
    for k in ("recall", "precision", "ndcg"):
        out["lg_test_" + k] = res[k]
    # trained-like weights (scores of a few units; SURVEY section 8 d) for ranking parity
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        m.user_embedding.weight.copy_(torch.randn(m.user_embedding.weight.shape, generator=g) * 0.4)
        m.item_embedding.weight.copy_(torch.randn(m.item_embedding.weight.shape, generator=g) * 0.4)
    out["lg_user_wT"], out["lg_item_wT"] = m.user_embedding.weight.detach().numpy().copy(), m.item_embedding.weight.detach().numpy().copy()
    res = ref_test.Test(data, m, dev, cfg)
    for k in ("recall", "precision", "ndcg"):
        out["lgT_test_" + k] = res[k]
    with torch.no_grad():
        users = list(data.test_dict.keys())[:50]
        rating = m.get_rating_for_test(torch.Tensor(users).long())
        out["lgT_rating50"] = rating.numpy().copy()
        fu, fi = m.aggregate()
        out["lgT_fu"], out["lgT_fi"] = fu.numpy().copy(), fi.numpy().copy()

    # ---- SimGCL / XSimGCL with captured noise
    for kind, extra in (("SimGCL", dict(ssl_lambda=0.5, temperature=0.2, epsilon=0.05)),
                        ("XSimGCL", dict(ssl_lambda=0.2, temperature=0.15, epsilon=0.2, cl_layer=1)),
                        ("XSimGCL2", dict(ssl_lambda=0.2, temperature=0.15, epsilon=0.2, cl_layer=2))):
        mod = __import__("models." + kind.rstrip("2"), fromlist=["x"])
        cls = getattr(mod, kind.rstrip("2"))
        c2 = dict(cfg)
        c2.update({k: str(v) for k, v in extra.items()})
        ref_tools.set_seed(2024)
        mm = cls(c2, data, dev)
        rec = RecordingRand()
        torch.rand_like = rec
        try:
            ll = mm(bu, bp, bn)
        finally:
            torch.rand_like = rec.orig
        tot = sum(ll)
        tot.backward()
        p = kind.lower()
        out[p + "_noise"] = np.stack([d.numpy() for d in rec.draws])
        out[p + "_loss"] = np.array([l.item() for l in ll])
        out[p + "_gu"] = mm.user_embedding.weight.grad.numpy().copy()
        out[p + "_gi"] = mm.item_embedding.weight.grad.numpy().copy()
        out[p + "_user_w0"] = mm.user_embedding.weight.detach().numpy().copy()
        out[p + "_item_w0"] = mm.item_embedding.weight.detach().numpy().copy()
        with torch.no_grad():
            fu, fi = mm.aggregate(perturbed=False)
        out[p + "_fu"], out[p + "_fi"] = fu.numpy().copy(), fi.numpy().copy()

    # ---- NGCF with recorded dropout masks
    from models.NGCF import NGCF
    c3 = dict(cfg)
    c3.update({"mess_dropout": "True", "mess_drop_prob": "[0.1, 0.1, 0.1]", "node_dropout": "False",
               "node_drop_prob": "0.1", "layer_size": "[64, 64, 64]", "learn_rate": "0.0001"})
    ref_tools.set_seed(2024)
    mn = NGCF(c3, data, dev)
    for k, v in mn.weight_dict.items():
        out["ngcf_" + k] = v.detach().numpy().copy()
    out["ngcf_user_w0"] = mn.user_embedding.weight.detach().numpy().copy()
    out["ngcf_item_w0"] = mn.item_embedding.weight.detach().numpy().copy()
    orig = torch.nn.Dropout
    RecordingDropout.masks = []
    torch.nn.Dropout = RecordingDropout
    try:
        ll = mn(bu, bp, bn)
    finally:
        torch.nn.Dropout = orig
    sum(ll).backward()
    out["ngcf_masks"] = np.stack([k.numpy() for k in RecordingDropout.masks])
    out["ngcf_loss"] = np.array([l.item() for l in ll])
    out["ngcf_gu"] = mn.user_embedding.weight.grad.numpy().copy()
    out["ngcf_gi"] = mn.item_embedding.weight.grad.numpy().copy()
    for k, v in mn.weight_dict.items():
        out["ngcf_g_" + k] = v.grad.numpy().copy()

    # ---- MFBPR step (no graph)
    from models.MFBPR import MFBPR
    ref_tools.set_seed(2024)
    mf = MFBPR(cfg, data, dev)
    ll = mf(bu, bp, bn)
    sum(ll).backward()
    out["mf_loss"] = np.array([l.item() for l in ll])
    out["mf_gu"] = mf.user_embedding.weight.grad.numpy().copy()


def run_functional(out):
    g = torch.Generator().manual_seed(11)
    a, b, c = (torch.randn(37, 64, generator=g) for _ in range(3))
    out["fn_a"], out["fn_b"], out["fn_c"] = a.numpy(), b.numpy(), c.numpy()
    out["fn_bpr"] = ref_losses.get_bpr_loss(a, b, c).item()
    out["fn_reg"] = ref_losses.get_reg_loss(a, b, c).item()
    out["fn_nce"] = ref_losses.get_InfoNCE_loss(a, b, 0.2).item()
    # metrics known answers
    truth = [[1, 2, 3], [7], [4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26]]
    pred = np.array([[1, 9, 2, 8, 3], [0, 1, 2, 3, 4], [4, 5, 6, 8, 9]])
    r = ref_metrics.get_label(truth, pred)
    out["mt_r"] = r
    out["mt_vals"] = np.array([[ref_metrics.recall_at_k(r, k, truth), ref_metrics.precision_at_k(r, k, truth),
                                ref_metrics.ndcg_at_k(r, k, truth)] for k in (3, 5)])


def main():
    root = tempfile.mkdtemp(prefix="idgrec_golden_")
    try:
        # dataset 1: clean tiny graph
        g = datagen.gen_graph("tiny", seed=2024)
        datagen.write_dataset(root, "tiny", g)
        # dataset 2: duplicates in train.txt (weight-2 edges), a user with an empty line,
        # an item id that only occurs in test (num_items comes from both files)
        g2 = datagen.gen_graph((40, 60, 500, 140), seed=5)
        d2 = datagen.write_dataset(root, "quirks", g2)
        lines = open(d2 + "/train.txt").read().splitlines()
        lines[3] = lines[3] + " " + lines[3].split(" ")[1]            # duplicate pair
        lines[7] = lines[7] + " " + " ".join(lines[7].split(" ")[1:3])  # two more duplicates
        lines.append("40")                                             # user with no item
        open(d2 + "/train.txt", "w").write("\n".join(lines) + "\n")
        tl = open(d2 + "/test.txt").read().splitlines()
        tl[0] = tl[0] + " 63"
        open(d2 + "/test.txt", "w").write("\n".join(tl) + "\n")

        for name in ("tiny", "quirks"):
            out = {}
            data, cfg = run_dataset(name, root, out)
            shutil.copy(root + "/" + name + "/train.txt", os.path.join(OUT, name + "_train.txt"))
            shutil.copy(root + "/" + name + "/test.txt", os.path.join(OUT, name + "_test.txt"))
            if name == "tiny":
                run_models(name, data, cfg, out)
                run_functional(out)
            np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
            print(name, "->", len(out), "arrays")
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
