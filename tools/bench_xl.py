"""Scale-up benchmark (BASELINE.json configs[4]): LightGCN 3-layer d=64 on a synthetic 1M x 1M / 100M-edge graph,
row-partitioned over 1/2/4/8 B200s.  Reports the train-step time (max over ranks, CUDA events), the propagation
layer rate against the HBM roofline, and user-sharded full-ranking throughput.

    python -m torch.distributed.run --nproc-per-node G tools/bench_xl.py [--users 1000000 --items 1000000 --edges 100000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "id-grec_b200"), REPO):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=1000000)
    ap.add_argument("--items", type=int, default=1000000)
    ap.add_argument("--edges", type=int, default=100000000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--eval-users", type=int, default=16384, help="test users ranked per rank")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--breakdown", action="store_true")
    ap.add_argument("--closure", default="auto", choices=["auto", "0", "1"])
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from idgrec import _lib, datagen, ops
    from idgrec.dist import DistFusedTrainer
    from idgrec.graph import Graph, build_norm_adjacency
    U, I, E, d, K, B = args.users, args.items, args.edges, 64, 3, args.batch
    N = U + I
    t0 = time.time()
    eu, ei = datagen.gen_edges_device(U, I, E, 2024, dev)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    csr = build_norm_adjacency(eu, ei, U, I, device=dev)
    torch.cuda.synchronize()
    t_csr = time.time() - t0
    nnz = csr.nnz
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    bound = (6.0 / (U + d)) ** 0.5
    table = (torch.rand(N, d, generator=gen, device=dev) * 2 - 1) * bound
    ft = DistFusedTrainer("LightGCN", csr, table, U, K, 1e-4, 1e-3, rank, world, max_batch=B, use_cuda_graph=not args.no_graph,
                           closure_restrict={"0": False, "1": True}.get(args.closure, "auto"))
    del table
    nb = args.steps + args.warmup
    sel = torch.randint(0, E, (nb, B), generator=gen, device=dev)
    negs = torch.randint(0, I, (nb, B), generator=gen, device=dev)
    bu, bp = eu[sel], ei[sel]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for s in range(args.warmup):
        ft.step(bu[s], bp[s], negs[s])
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(args.warmup, nb):
        ft.step(bu[s], bp[s], negs[s])
    b.record()
    sync()
    step_ms = a.elapsed_time(b) / args.steps
    breakdown = ft.profile_steps([(bu[s], bp[s], negs[s]) for s in range(min(8, nb))]) if args.breakdown else None
    sync()
    # one local propagation layer (local rows, peer stores on) timed alone
    evs = []
    for _ in range(8):
        x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x.record()
        ft.local.spmm_layer(ft.E0, Y=ft.W[0])
        y.record()
        ft.slab.barrier()
        evs.append((x, y))
    sync()
    layer_ms = float(np.median([x.elapsed_time(y) for x, y in evs]))
    # evaluation shard: args.eval_users users of this rank's shard against all items, mask = train positives
    F = ft.final_embeddings()
    ip = csr.indptr[:U + 1].contiguous()
    mask_idx = (csr.indices[: int(ip[-1].item())] - U).contiguous()
    from idgrec.dist import shard_range
    s0, s1 = shard_range(U, rank, world)
    nu = min(args.eval_users, s1 - s0)
    users = torch.arange(s0, s0 + nu, device=dev)
    ws = torch.empty(int(_lib.lib().idg_eval_workspace_bytes(nu, I, d, 20)), dtype=torch.uint8, device=dev)
    ops.eval_topk(F[:U], F[U:], users[:256], ip, mask_idx, 20, ws=ws)
    sync()
    a.record()
    ops.eval_topk(F[:U], F[U:], users, ip, mask_idx, 20, ws=ws)
    b.record()
    sync()
    eval_ms = a.elapsed_time(b)
    t = torch.tensor([step_ms, layer_ms, eval_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, layer_ms, eval_ms = t.tolist()
    if rank == 0:
        try:
            hbm = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm = 6650.0
        loc_rows, loc_nnz = ft.b1 - ft.b0, ft.local.nnz
        alg = 4 * (loc_rows + 1) + 8 * loc_nnz + 4 * N * d / world + 4 * loc_rows * d   # CSR slice + share of X + local Y
        gat = 8 * loc_nnz + 4 * loc_nnz * d + 4 * loc_rows * d
        print(json.dumps({"workload": "LightGCN 3-layer d=64 train step, %d users / %d items / %d edges (nnz %d), batch %d" % (U, I, E, nnz, B),
                          "n_gpus": world, "ms_per_train_step": step_ms, "steps": args.steps, "warmup": args.warmup,
                          "layer_ms_local_rows": layer_ms, "layer_alg_GBs_per_gpu": alg / layer_ms / 1e6, "layer_gather_GBs_per_gpu": gat / layer_ms / 1e6,
                          "layer_gather_frac_of_hbm": gat / layer_ms / 1e6 / hbm, "hbm_peak_GBs": hbm,
                          "eval_users_per_s_total": nu * world / eval_ms * 1e3, "eval_ms": eval_ms, "eval_users_per_rank": nu,
                          "gen_s": t_gen, "csr_build_s": t_csr, "bounds": ft.bounds, "cuda_graph": not args.no_graph, "breakdown_ms": breakdown, "closure_restrict": ft.use_closure, "slab_backend": ft.slab.backend, "multicast": ft.slab.multicast}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
