"""Where does Test() spend its time?  CUDA-event breakdown of the evaluator's phases on the amazon-book shape, and the
per-shard ranking time / flagged-user count on the XL graph (the r1 record showed eval_ms 9.7 -> 178 from 1 to 8 GPUs).

    python tools/diag_eval.py [--xl]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "id-grec_b200"), REPO):
    sys.path.insert(0, p)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def flagged(ws):
    return int(ws[256:260].view(torch.int32).item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--xl", action="store_true")
    ap.add_argument("--shape", default="amazon-book")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    from idgrec import _lib, datagen, ops
    from idgrec.graph import Graph, build_norm_adjacency
    l = _lib.lib()
    out = {}
    if not args.xl:
        from utility.utility_data.data_loader import Data
        import utility.utility_function.tools as tools
        import utility.utility_train.batch_test as batch_test
        from models.LightGCN import LightGCN
        cfg = {"embedding_size": "64", "batch_size": "1024", "test_batch_size": "1024", "learn_rate": "0.001", "reg_lambda": "0.0001",
               "GCN_layer": "3", "top_K": "[10, 20]", "sparsity_test": "0", "dataset": "synthetic"}
        g = datagen.gen_graph(args.shape)
        data = Data.from_arrays(g.num_users, g.num_items, g.train_user, g.train_item, g.test_user, g.test_item, cfg)
        tools.set_seed(2024)
        model = LightGCN(cfg, data, dev).to(dev)
        import utility.utility_train.trainer as trainer
        for scale_name, scale in (("xavier", None), ("trained 1 epoch", "train"), ("trained 3 epochs", "train2"), ("trained-like 0.4", 0.4)):
            if scale in ("train", "train2"):
                ft = model.fused_trainer(1e-3, 1024)
                for _ in range(1 if scale == "train" else 2):
                    users_, pos_, neg_ = trainer.sample_epoch(data, dev)
                    for s in range(0, len(users_), 1024):
                        ft.step(users_[s:s + 1024], pos_[s:s + 1024], neg_[s:s + 1024])
                torch.cuda.synchronize()
            elif scale is not None:
                with torch.no_grad():
                    model._table.normal_(0, scale)
            for _ in range(3):
                batch_test.Test(data, model, dev, cfg)
            torch.cuda.synchronize()
            cache = data.device_cache(dev)
            rec = {}
            t0 = time.perf_counter()
            e0 = ev()
            fu, fi = model.final_embeddings()
            e1 = ev()
            users = cache["test_users"]
            ws = torch.empty(int(l.idg_eval_workspace_bytes(len(users), data.num_items, 64, 20)), dtype=torch.uint8, device=dev)
            e2 = ev()
            ids = ops.eval_topk(fu, fi, users, cache["mask_indptr"], cache["mask_indices"], 20, ws=ws)
            e3 = ev()
            sums = ops.eval_metric_sums(ids, users, cache["test_indptr"], cache["test_indices"], [10, 20])
            e4 = ev()
            s = sums.cpu()
            torch.cuda.synchronize()
            rec["wall_ms"] = (time.perf_counter() - t0) * 1e3
            rec["propagate_ms"] = e0.elapsed_time(e1)
            rec["alloc_ms"] = e1.elapsed_time(e2)
            rec["topk_ms"] = e2.elapsed_time(e3)
            rec["metrics_ms"] = e3.elapsed_time(e4)
            rec["flagged_users"] = flagged(ws)
            nrm = fi.norm(dim=1)
            rec["item_norm_max_median_p99"] = [float(nrm.max()), float(nrm.median()), float(torch.quantile(nrm, 0.99))]
            cc = ws[512 + ((4 * len(users) + 255) // 256 * 256):][: 4 * len(users)].view(torch.int32)   # cand_cnt follows flag_list
            rec["cand_cnt_mean_max"] = [float(cc.float().mean()), int(cc.max())]
            rec["n_users"] = len(users)
            a, b = ev(), None
            batch_test.Test(data, model, dev, cfg)
            b = ev()
            torch.cuda.synchronize()
            rec["Test_ms_events"] = a.elapsed_time(b)
            out[scale_name] = rec
    else:
        U = I = 1000000
        E = 100000000
        eu, ei = datagen.gen_edges_device(U, I, E, 2024, dev)
        csr = build_norm_adjacency(eu, ei, U, I, device=dev)
        G = Graph(csr)
        gen = torch.Generator(device=dev)
        gen.manual_seed(7)
        bound = (6.0 / (U + 64)) ** 0.5
        table = (torch.rand(U + I, 64, generator=gen, device=dev) * 2 - 1) * bound
        F = G.propagate_fwd(table, 3, True)
        ip = csr.indptr[:U + 1].contiguous()
        mask_idx = (csr.indices[: int(ip[-1].item())] - U).contiguous()
        nu = 16384
        ws = torch.empty(int(l.idg_eval_workspace_bytes(nu, I, 64, 20)), dtype=torch.uint8, device=dev)
        rows = []
        for s0 in range(0, U, 125000):
            users = torch.arange(s0, s0 + nu, device=dev)
            ops.eval_topk(F[:U], F[U:], users[:256], ip, mask_idx, 20, ws=ws)
            torch.cuda.synchronize()
            a = ev()
            ops.eval_topk(F[:U], F[U:], users, ip, mask_idx, 20, ws=ws)
            b = ev()
            torch.cuda.synchronize()
            rows.append({"first_user": s0, "ms": a.elapsed_time(b), "flagged": flagged(ws)})
        out["xl_shards"] = rows
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
