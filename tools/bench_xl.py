"""Scale-up benchmark (BASELINE.json configs[4]): LightGCN 3-layer d=64 on a synthetic 1M x 1M / 100M-edge graph,
row-partitioned over 1/2/4/8 B200s.  Reports the train-step time (max over ranks, CUDA events), the same run's
one-GPU step time (so the speed-up is measured inside one job), the propagation layer rate against the HBM
roofline, the bytes every GPU receives per exchanged layer, and the user-sharded full ranking of all 1M users.

    python -m torch.distributed.run --nproc-per-node G tools/bench_xl.py [--users 1000000 --items 1000000 --edges 100000000]

``run()`` is what bench.py calls for the ``xl`` sub-record of its JSON line.
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
    if p not in sys.path:
        sys.path.insert(0, p)


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def run(rank, world, dev, users=1000000, items=1000000, edges=100000000, steps=20, warmup=5, batch=1024, eval_users=None,
        use_graph=True, breakdown=False, closure="auto", one_gpu_reference=True, eval_chunk=65536):
    """Returns the record on rank 0 (None elsewhere).  eval_users: users ranked per rank (None = the rank's whole shard,
    i.e. all `users` are ranked once across the job)."""
    from idgrec import _lib, datagen, ops
    from idgrec.dist import DistFusedTrainer, shard_range
    from idgrec.graph import Graph, build_norm_adjacency
    U, I, E, d, K, B = users, items, edges, 64, 3, batch
    N = U + I

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

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
    nb = steps + warmup
    sel = torch.randint(0, E, (nb, B), generator=gen, device=dev)
    negs = torch.randint(0, I, (nb, B), generator=gen, device=dev)
    bu, bp = eu[sel], ei[sel]
    del eu, ei, sel
    full = Graph(csr)
    clo = {"0": False, "1": True}.get(str(closure), "auto")

    def timed_steps(ft):
        for s in range(warmup):
            ft.step(bu[s], bp[s], negs[s])
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(warmup, nb):
            ft.step(bu[s], bp[s], negs[s])
        b.record()
        sync()
        return a.elapsed_time(b) / steps

    # ---- the same job's one-GPU figure (every rank runs it on its own GPU; no communication) ----
    n1_ms = None
    if world > 1 and one_gpu_reference:
        ft1 = DistFusedTrainer("LightGCN", csr, table, U, K, 1e-4, 1e-3, 0, 1, max_batch=B, use_cuda_graph=use_graph, full_graph=full,
                               closure_restrict=clo)
        n1_ms = timed_steps(ft1)
        del ft1
        torch.cuda.empty_cache()

    ft = DistFusedTrainer("LightGCN", csr, table, U, K, 1e-4, 1e-3, rank, world, max_batch=B, use_cuda_graph=use_graph, full_graph=full,
                          closure_restrict=clo)
    del table
    step_ms = timed_steps(ft)
    phases = ft.profile_steps([(bu[s], bp[s], negs[s]) for s in range(min(8, nb))]) if breakdown else None
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
    ft.slab.status()

    # ---- user-sharded full ranking: every rank ranks its shard of ALL users against all items, mask = train positives ----
    F = ft.final_embeddings()
    ip = csr.indptr[:U + 1].contiguous()
    mask_idx = (csr.indices[: int(ip[-1].item())] - U).contiguous()
    s0, s1 = shard_range(U, rank, world)
    if eval_users is not None:
        s1 = min(s1, s0 + eval_users)
    nu = s1 - s0
    chunk = min(eval_chunk, max(nu, 1))
    ws = torch.empty(int(_lib.lib().idg_eval_workspace_bytes(chunk, I, d, 20)), dtype=torch.uint8, device=dev)
    ops.eval_topk(F[:U], F[U:], torch.arange(s0, s0 + min(256, nu), device=dev), ip, mask_idx, 20, ws=ws)
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    marks = [a]
    for c0 in range(s0, s1, chunk):
        ops.eval_topk(F[:U], F[U:], torch.arange(c0, min(s1, c0 + chunk), device=dev), ip, mask_idx, 20, ws=ws)
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append(e)
    b.record()
    sync()
    eval_ms = a.elapsed_time(b)
    chunk_ms = [round(x.elapsed_time(y), 3) for x, y in zip(marks[:-1], marks[1:])]
    vals = [step_ms, layer_ms, eval_ms, n1_ms if n1_ms is not None else 0.0]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    nus = torch.tensor([nu], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nus)
    step_ms, layer_ms, eval_ms, n1_max = t.tolist()
    rec = None
    if rank == 0:
        hbm, hbm_src = _hbm_peak()
        loc_rows, loc_nnz = ft.b1 - ft.b0, ft.local.nnz
        alg = 4 * (loc_rows + 1) + 8 * loc_nnz + 4 * N * d / world + 4 * loc_rows * d   # CSR slice + share of X + local Y
        gat = 8 * loc_nnz + 4 * loc_nnz * d + 4 * loc_rows * d
        recv = (N - loc_rows) * d * 4 if world > 1 else 0
        rec = {"workload": "LightGCN 3-layer d=64 train step, %d users / %d items / %d edges (nnz %d), batch %d, row-partitioned x%d" % (U, I, E, nnz, B, world),
               "n_gpus": world, "ms_per_train_step": step_ms, "steps": steps, "warmup": warmup,
               "one_gpu_ms_per_train_step_same_job": (n1_max if world > 1 and n1_ms is not None else step_ms),
               "speedup_vs_one_gpu": ((n1_max / step_ms) if world > 1 and n1_ms is not None else 1.0),
               "layer_ms_local_rows": layer_ms, "layer_alg_bytes_per_gpu": alg, "layer_alg_GBs_per_gpu": alg / layer_ms / 1e6,
               "layer_frac_of_hbm_compulsory": alg / layer_ms / 1e6 / hbm,
               "layer_gather_GBs_per_gpu": gat / layer_ms / 1e6, "layer_gather_frac_of_hbm": gat / layer_ms / 1e6 / hbm,
               "hbm_peak_GBs": hbm, "hbm_peak_source": hbm_src,
               "bytes_received_per_exchanged_layer_per_gpu": recv, "exchanged_layers_per_step": 4 if world > 1 else 0,
               "nvlink_floor_ms_per_layer": recv / 770e6 if world > 1 else 0.0,
               "eval_users_total": int(nus.item()), "eval_ms": eval_ms, "eval_chunk_users": chunk, "eval_chunk_ms_rank0": chunk_ms, "eval_users_per_s_total": int(nus.item()) / eval_ms * 1e3,
               "gen_s": t_gen, "csr_build_s": t_csr, "bounds": ft.bounds, "cuda_graph": bool(use_graph), "breakdown_ms": phases,
               "closure_restrict": ft.use_closure, "exchange": ("chunked x%d, %d push CTAs" % (len(ft.chunks), ft.push_ctas)) if ft.chunked else "fused epilogue stores", "slab_backend": ft.slab.backend, "multicast": ft.slab.multicast,
               "one_gpu_schedule_note": "the one-GPU figure runs the class-split schedule of csrc/spmm.cu:classify_rows (all user rows, then all item rows: the L2 holds one "
                                        "256 MB half of the table at a time; 18.9 -> 16.1 ms per step); a rank of a multi-GPU job holds (almost) one class already and gains nothing, "
                                        "so speedup_vs_one_gpu is smaller than before that change (5.6x -> ~4.8x at 8 GPUs) although no step got slower"}
    del ft, F, full, csr, ws
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=1000000)
    ap.add_argument("--items", type=int, default=1000000)
    ap.add_argument("--edges", type=int, default=100000000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--eval-users", type=int, default=0, help="test users ranked per rank (0 = the whole shard)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--breakdown", action="store_true")
    ap.add_argument("--closure", default="auto", choices=["auto", "0", "1"])
    ap.add_argument("--no-one-gpu-ref", action="store_true", help="skip the same-job one-GPU reference run")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rec = run(rank, world, dev, args.users, args.items, args.edges, args.steps, args.warmup, args.batch, args.eval_users or None,
              not args.no_graph, args.breakdown, args.closure, not args.no_one_gpu_ref)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
