"""torchrun --nproc-per-node G tools/dist_check.py : the row-partitioned G-GPU trainer must reproduce the
1-GPU fused trainer bit for bit (weights after several steps) and the sharded evaluation must give the same
metrics.  Used by tests/test_gpu_dist.py and by hand on the GPU box."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "id-grec_b200"), REPO):
    sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from idgrec import datagen
    from idgrec.engine import FusedTrainer
    from utility.utility_data.data_loader import Data
    import utility.utility_function.tools as tools
    import utility.utility_train.batch_test as batch_test
    import importlib
    shape = sys.argv[1] if len(sys.argv) > 1 else "small"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    kind = sys.argv[3] if len(sys.argv) > 3 else "LightGCN"          # LightGCN | SimGCL | XSimGCL | XSimGCL3 (cl_layer = 3)
    extra = {"LightGCN": {}, "SimGCL": {"ssl_lambda": "0.5", "temperature": "0.2", "epsilon": "0.05"},
             "XSimGCL": {"ssl_lambda": "0.2", "temperature": "0.15", "epsilon": "0.2", "cl_layer": "1"},
             "XSimGCL3": {"ssl_lambda": "0.2", "temperature": "0.15", "epsilon": "0.2", "cl_layer": "3"}}[kind]
    kind = kind.rstrip("3")
    Model = getattr(importlib.import_module("models." + kind), kind)
    cfg = {"embedding_size": "64", "batch_size": "1024", "test_batch_size": "1024", "learn_rate": "0.001", "reg_lambda": "0.0001",
           "GCN_layer": "3", "top_K": "[10, 20]", "sparsity_test": "0", "dataset": "synthetic", "cuda_graph": os.environ.get("IDG_GRAPH", "1"),
           "closure_restrict": os.environ.get("IDG_CLOSURE", "auto")}
    cfg.update(extra)
    g = datagen.gen_graph(shape)
    data = Data.from_arrays(g.num_users, g.num_items, g.train_user, g.train_item, g.test_user, g.test_item, cfg)
    tools.set_seed(2024)
    model = Model(cfg, data, dev)
    model.to(dev)
    w0 = model._table.clone()
    ft = model.fused_trainer(1e-3, 1024)
    assert type(ft).__name__ == "DistFusedTrainer", type(ft)
    rng = np.random.default_rng(5)
    batches = []
    for _ in range(steps):
        e = rng.integers(0, len(g.train_user), 1024)
        batches.append(tuple(torch.from_numpy(a).to(dev) for a in (g.train_user[e], g.train_item[e], rng.integers(0, g.num_items, 1024))))
    # the perturbed views: the same injected noise on every rank and in the single-GPU reference (one [K,N,d] block per
    # view and step; the trainers' own draws come from the device generator and match across ranks only by seed)
    n_views = {"LightGCN": 0, "SimGCL": 2, "XSimGCL": 1}[kind]
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    noises = [[torch.rand(3, data.num_nodes, 64, generator=gen, device=dev) for _ in range(n_views)] for _ in range(steps)]
    nz_buf = [torch.empty(3, data.num_nodes, 64, device=dev) for _ in range(n_views)]   # stable pointers for the captured step
    ft.injected_noise = nz_buf if n_views else None
    losses = []
    for b, nz in zip(batches, noises):
        for dst, src in zip(nz_buf, nz):
            dst.copy_(src)
        losses.append(ft.step(*b).clone())
    torch.cuda.synchronize()
    res = batch_test.Test(data, model, dev, cfg)
    dist.barrier()
    ok = True
    if rank == 0:
        # single-GPU reference on the same batches
        ref = FusedTrainer(kind, model.Graph, w0.clone(), data.num_users, 3, 1e-4, 1e-3, max_batch=1024, use_cuda_graph=False,
                           ssl_lambda=float(cfg.get("ssl_lambda", 0.0)), temperature=float(cfg.get("temperature", 0.2)),
                           eps=float(cfg.get("epsilon", 0.0)), cl_layer=int(cfg.get("cl_layer", 1)))
        for b, l, nz in zip(batches, losses, noises):
            ref.injected_noise = nz if n_views else None
            lr = ref.step(*b)
            ok &= bool(torch.equal(lr, l))
        same = torch.equal(ref.E0, ft.E0)
        print("losses identical:", ok, "| tables bit-identical:", same, "| max abs diff:", float((ref.E0 - ft.E0).abs().max()))
        ok &= same
    # all ranks hold the same table
    t = ft.E0.clone()
    dist.broadcast(t, 0)
    ok &= bool(torch.equal(t, ft.E0))
    # sharded evaluation == single-GPU evaluation: every rank ranks its user shard (what Test() does under world > 1);
    # rank 0 also ranks ALL users from the single-GPU trainer's table.  ids must be bit-identical, metric sums equal to
    # float64 summation order.
    from idgrec import ops
    from idgrec.dist import shard_range
    cache = data.device_cache(dev)
    all_users = cache["test_users"]
    s0, s1 = shard_range(len(all_users), rank, world)
    topK = eval(cfg["top_K"])
    my_ids, _ = batch_test.rank_all(data, model, dev, max(topK), users=all_users[s0:s1].contiguous())
    per = (len(all_users) + world - 1) // world
    pad = torch.full((per, max(topK)), -1, dtype=torch.int64, device=dev)
    pad[: s1 - s0] = my_ids
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    if rank == 0:
        class SingleGpuView:   # the 1-GPU trainer's table behind the evaluator's model contract
            def final_embeddings(self_inner):
                F = model.Graph.propagate_fwd(ref.E0, 3, kind == "LightGCN")
                return F[:data.num_users], F[data.num_users:]
        ids1, users1 = batch_test.rank_all(data, SingleGpuView(), dev, max(topK))
        sharded = torch.cat([gathered[r][: shard_range(len(all_users), r, world)[1] - shard_range(len(all_users), r, world)[0]] for r in range(world)])
        ids_same = bool(torch.equal(sharded, ids1))
        sums1 = ops.eval_metric_sums(ids1, users1, cache["test_indptr"], cache["test_indices"], topK).cpu().numpy() / float(len(all_users))
        met_same = bool(np.allclose(sums1[:, 0], res["recall"], rtol=1e-12, atol=0) and np.allclose(sums1[:, 2], res["ndcg"], rtol=1e-12, atol=0)
                        and np.allclose(sums1[:, 1], res["precision"], rtol=1e-12, atol=0))
        print("sharded top-K ids == single-GPU ids:", ids_same, "| sharded Test() metrics == single-GPU metrics:", met_same)
        ok &= ids_same and met_same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("closure", ft.use_closure, "recall@20", res["recall"][1], "ndcg@20", res["ndcg"][1], "bounds", ft.bounds, "slab", ft.slab.backend, "multicast", ft.slab.multicast)
        print("DIST_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
