"""Generates tests/golden/next.npz by running the UNMODIFIED reference modules on the tiny / quirks
datasets already committed next to it (SURVEY.md section 8 f rows: sparsity_test evaluator, SGL and
the LightGCN-backbone models that only add a batch x batch loss).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_next.py
Nothing from the reference is copied: it is imported from where it lies, only its outputs are stored.
"""
import os
import random
import shutil
import sys
import tempfile

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")

for m in [k for k in sys.modules if k.split(".")[0] in ("utility", "models", "Parser")]:
    del sys.modules[m]
sys.path.insert(0, REF)
import utility.utility_data.data_loader as ref_loader  # noqa: E402
import utility.utility_function.tools as ref_tools  # noqa: E402
import utility.utility_function.losses as ref_losses  # noqa: E402
import utility.utility_train.batch_test as ref_test  # noqa: E402

assert ref_loader.__file__.startswith(REF), ref_loader.__file__


def ragged(list_of_arrays):
    ptr = np.cumsum([0] + [len(a) for a in list_of_arrays]).astype(np.int64)
    flat = np.concatenate([np.asarray(a, dtype=np.int64) for a in list_of_arrays]) if ptr[-1] else np.zeros(0, np.int64)
    return ptr, flat


def base_config(name, root, **kw):
    c = {"dataset_path": root + "/", "dataset": name, "top_K": "[10, 20]", "training_epochs": "2",
         "early_stopping": "10", "interval": "1", "embedding_size": "64", "batch_size": "256",
         "test_batch_size": "50", "learn_rate": "0.001", "reg_lambda": "0.0001", "GCN_layer": "3",
         "sparsity_test": "0"}
    c.update({k: str(v) for k, v in kw.items()})
    return c


def stage(root, name):
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    shutil.copy(os.path.join(OUT, name + "_train.txt"), os.path.join(d, "train.txt"))
    shutil.copy(os.path.join(OUT, name + "_test.txt"), os.path.join(d, "test.txt"))
    return d


def grads(model):
    return model.user_embedding.weight.grad.numpy().copy(), model.item_embedding.weight.grad.numpy().copy()


def main():
    tiny = np.load(os.path.join(OUT, "tiny.npz"))
    root = tempfile.mkdtemp(prefix="idgrec_golden_next_")
    out = {}
    dev = torch.device("cpu")
    try:
        # ---------------- sparsity split + sparsity_test (data_loader.py:161-204, batch_test.py:110-170)
        for name in ("tiny", "quirks"):
            stage(root, name)
            cfg = base_config(name, root, sparsity_test=1, top_K="[20, 40]", test_batch_size=37)
            data = ref_loader.Data(root + "/" + name, cfg)
            out["split_%s_ptr" % name], out["split_%s_flat" % name] = ragged(data.split_test_dict)
            out["split_%s_state" % name] = np.array(data.split_state)
            if name == "tiny":
                from models.LightGCN import LightGCN
                ref_tools.set_seed(2024)
                m = LightGCN(cfg, data, dev)
                with torch.no_grad():
                    m.user_embedding.weight.copy_(torch.from_numpy(tiny["lg_user_wT"]))
                    m.item_embedding.weight.copy_(torch.from_numpy(tiny["lg_item_wT"]))
                try:
                    res = ref_test.sparsity_test(data, m, dev, cfg)
                    out["sparsity_ok"] = True
                    for k in ("recall", "precision", "ndcg"):
                        out["sparsity_" + k] = np.stack([r[k] for r in res])
                except AssertionError:
                    out["sparsity_ok"] = False
                # plain Test() with the wider top-K list on the same weights
                cfg0 = dict(cfg, sparsity_test="0")
                res = ref_test.Test(data, m, dev, cfg0)
                for k in ("recall", "precision", "ndcg"):
                    out["test2040_" + k] = res[k]

        # ---------------- loss-only models on the LightGCN backbone, one batch, losses + gradients
        cfg = base_config("tiny", root)
        data = ref_loader.Data(root + "/tiny", cfg)
        B = 256
        samples = tiny["sample_ep0"][tiny["perm_ep0"]]
        bu, bp, bn = (torch.from_numpy(samples[:B, j].copy()).long() for j in range(3))
        out["batch"] = samples[:B]
        specs = {
            "LightCCF": dict(ssl_lambda=5.0, temperature=0.22),
            "LightCSCF": dict(lambda_reg=0.0001, lambda_gamma=1.0, lambda_margin=0.7, temperature=0.2),
            "SCCF": dict(temperature=0.1),
            "DirectAU": dict(gamma=2.0),
        }
        for kind, extra in specs.items():
            for enc in ("LightGCN", "MF"):
                mod = __import__("models." + kind, fromlist=["x"])
                c2 = dict(cfg, encoder=enc)
                c2.update({k: str(v) for k, v in extra.items()})
                ref_tools.set_seed(2024)
                mm = getattr(mod, kind)(c2, data, dev)
                assert np.array_equal(mm.user_embedding.weight.detach().numpy(), tiny["lg_user_w0"])
                ll = mm(bu, bp, bn)
                sum(ll).backward()
                p = "%s_%s" % (kind.lower(), enc.lower())
                out[p + "_loss"] = np.array([l.item() for l in ll])
                out[p + "_gu"], out[p + "_gi"] = grads(mm)

        # LightCSCF with a margin low enough that the relu branch is live for many pairs
        from models.LightCSCF import LightCSCF
        c2 = dict(cfg, encoder="LightGCN", lambda_reg="0.0001", lambda_gamma="1.0", lambda_margin="0.05", temperature="0.2")
        ref_tools.set_seed(2024)
        mm = LightCSCF(c2, data, dev)
        with torch.no_grad():
            mm.user_embedding.weight.copy_(torch.from_numpy(tiny["lg_user_wT"]))
            mm.item_embedding.weight.copy_(torch.from_numpy(tiny["lg_item_wT"]))
        ll = mm(bu, bp, bn)
        sum(ll).backward()
        out["lightcscf_margin_loss"] = np.array([l.item() for l in ll])
        out["lightcscf_margin_gu"], out["lightcscf_margin_gi"] = grads(mm)

        # ---------------- SGL (edge dropping; python `random` is NOT seeded by set_seed -> record the kept edges)
        from models.SGL import SGL
        c2 = dict(cfg, ssl_lambda="0.1", ssl_ratio="0.1", aug_type="ed", temperature="0.2")
        ref_tools.set_seed(2024)
        mm = SGL(c2, data, dev)
        kept = []
        orig = random.sample

        def rec(pop, k):
            r = orig(pop, k)
            kept.append(np.array(r, dtype=np.int64))
            return r
        random.seed(99)
        random.sample = rec
        try:
            subs = [ref_tools.create_adj_mat(data.user_item_net, "ed", 0.1) for _ in range(2)]
        finally:
            random.sample = orig
        for j, s in enumerate(subs):
            s = s.tocsr().copy()
            s.sort_indices()
            out["sgl_keep%d" % j] = kept[j]
            out["sgl_sub%d_indptr" % j], out["sgl_sub%d_indices" % j], out["sgl_sub%d_data" % j] = s.indptr, s.indices, s.data
            assert s.data.dtype == np.float32
        g1, g2 = (ref_tools.convert_sp_mat_to_sp_tensor(s).to(dev) for s in subs)
        ll = mm(bu, bp, bn, g1, g2)
        sum(ll).backward()
        out["sgl_loss"] = np.array([l.item() for l in ll])
        out["sgl_gu"], out["sgl_gi"] = grads(mm)

        # ---------------- EGCF (item-only embeddings, R graph, tanh layers, three InfoNCE terms), both aggregate modes
        from models.EGCF import EGCF
        import utility.utility_data.data_graph as ref_graph
        Rm = ref_graph.sparse_adjacency_matrix_R(data).tocsr().copy()
        Rm.sort_indices()
        out["egcf_R_dtype"] = np.array(str(Rm.dtype))
        Rt = ref_tools.convert_sp_mat_to_sp_tensor(Rm).coalesce()     # what the model holds: fp32 values (tools.py:101)
        out["egcf_R_index"], out["egcf_R_value"] = Rt.indices().numpy(), Rt.values().numpy()
        for mode in ("parallel", "alternating"):
            c2 = dict(cfg, ssl_lambda="0.1", temperature="0.1", mode=mode)
            ref_tools.set_seed(2024)
            mm = EGCF(c2, data, dev)
            out["egcf_item_w0"] = mm.item_embedding.weight.detach().numpy().copy()
            with torch.no_grad():
                fu, fi = mm.parallel_aggregate() if mode == "parallel" else mm.alternating_aggregate()
            out["egcf_%s_fu" % mode], out["egcf_%s_fi" % mode] = fu.numpy().copy(), fi.numpy().copy()
            ll = mm(bu, bp, bn)
            sum(ll).backward()
            out["egcf_%s_loss" % mode] = np.array([l.item() for l in ll])
            out["egcf_%s_gi" % mode] = mm.item_embedding.weight.grad.numpy().copy()

        # ---------------- functional known answers for the added losses
        g = torch.Generator().manual_seed(11)
        a, b = (torch.randn(37, 64, generator=g) for _ in range(2))
        out["fn_a"], out["fn_b"] = a.numpy(), b.numpy()
        out["fn_align"] = ref_losses.get_align_loss(a, b).item()
        out["fn_uniform"] = ref_losses.get_uniform_loss(a).item()

        np.savez_compressed(os.path.join(OUT, "next.npz"), **out)
        print("next ->", len(out), "arrays;", "sparsity_ok =", out["sparsity_ok"])
        for k in sorted(out):
            if k.endswith("_loss"):
                print(k, out[k])
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
