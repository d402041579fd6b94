#!/usr/bin/env python
"""Headline benchmark: LightGCN 3-layer d=64 epoch seconds (train + full-ranking eval) on a
synthetic graph of the amazon-book shape (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this framework on N B200s
    python bench.py --impl reference --steps K --warmup W    # CPU port of the reference path, bounded sample

One "step" is ONE EPOCH of the hot path: every mini-batch of the epoch (propagate -> BPR ->
backward propagate -> Adam; 2,325 batches of 1,024 at the amazon-book shape) followed by one
full-ranking evaluation (propagate once, score all users x all items, mask train positives,
top-20, recall/ndcg).  `value` times that with the epoch's samples already on the device; `e2e`
times the same epoch through the public trainer API with host buffers (host negative sampling +
shuffle, pinned H2D copy of the samples, D2H of the loss and metric sums).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(REPO, "id-grec_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "LightGCN epoch s (train+full-rank eval)"
CFG = {"embedding_size": "64", "batch_size": "1024", "test_batch_size": "1024", "learn_rate": "0.001", "reg_lambda": "0.0001",
       "GCN_layer": "3", "top_K": "[10, 20]", "sparsity_test": "0", "dataset": "synthetic", "interval": "1",
       "training_epochs": "1", "early_stopping": "10"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", default="amazon-book")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batches", type=int, default=0, help="debug: cap the mini-batches per epoch (marks the line invalid)")
    ap.add_argument("--no-xl", action="store_true", help="skip the 1M x 1M / 100M-edge scale-up sub-record")
    ap.add_argument("--xl-steps", type=int, default=20)
    ap.add_argument("--no-configs", action="store_true", help="skip the yelp2018 / SimGCL / XSimGCL / NGCF sub-records (N=1 only)")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the reference-ops-on-the-GPU baseline (N=1 only)")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[j] for r in self.rows if len(r) >= 6 for j in range(4) if r[2 + j].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path (torch CPU sparse.mm / autograd / Adam /
# matmul + topk), bounded sample extrapolated to one epoch.  Test infrastructure used as checker
# and baseline only -- never on the product path.
# --------------------------------------------------------------------------------------------
def cpu_epoch_estimate(g, n_train_batches=2, n_eval_batches=1, threads=None, batch=1024, test_batch=1024):
    from oracle import ref_oracle as O
    import scipy.sparse as sp
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    U, I, E = g.num_users, g.num_items, len(g.train_user)
    net = sp.csr_matrix((np.ones(E), (g.train_user, g.train_item)), shape=(U, I))
    net.sort_indices()
    ip, ix, dt, _ = O.norm_adjacency(net)
    A = O.csr_to_torch_coo(ip, ix, dt, U + I)
    gen = torch.Generator().manual_seed(2024)
    om = O.OracleModel("LightGCN", A, O.xavier_uniform(U, 64, gen), O.xavier_uniform(I, 64, gen))
    rng = np.random.default_rng(1)
    e = rng.integers(0, E, batch)
    bu, bp, bn = g.train_user[e], g.train_item[e], rng.integers(0, I, batch)
    om.step(bu, bp, bn)  # warm-up
    t0 = time.perf_counter()
    for _ in range(n_train_batches):
        om.step(bu, bp, bn)
    t_batch = (time.perf_counter() - t0) / n_train_batches
    # evaluation as the reference does it: propagation re-run for every test batch (batch_test.py:59)
    users = np.unique(g.test_user)
    t0 = time.perf_counter()
    for b in range(n_eval_batches):
        fu, fi = om.final_embeddings()
        O.topk_reference_faithful(fu, fi, users[b * test_batch:(b + 1) * test_batch], net.indptr, net.indices, 20)
    t_eval = (time.perf_counter() - t0) / n_eval_batches
    nb = (E + batch - 1) // batch
    neb = (len(users) + test_batch - 1) // test_batch
    return {"epoch_s": t_batch * nb + t_eval * neb, "t_train_batch_s": t_batch, "t_eval_batch_s": t_eval, "train_batches": nb,
            "eval_batches": neb, "cores": threads,
            "sample": "%d train batches of %d (fwd+bwd+Adam) and %d eval batches of %d users (propagate+matmul+sigmoid+mask+topk), "
                      "extrapolated to %d + %d batches; sampling excluded" % (n_train_batches, batch, n_eval_batches, test_batch, nb, neb)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference is
    pure Python over torch/scipy and /root/reference does not exist on the GPU box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from idgrec import datagen
    g = datagen.gen_graph(args.shape)
    vals = []
    est = None
    for s in range(args.warmup + args.steps):
        est = cpu_epoch_estimate(g, n_train_batches=1, n_eval_batches=1)
        if s >= args.warmup:
            vals.append(est["epoch_s"])
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "s/epoch", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, g),
            "cpu_baseline": {"value": v, "unit": "s/epoch", "cores": est["cores"], "kind": "port", "sample": "per step: " + est["sample"]},
            "e2e": {"value": v, "unit": "s/epoch", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def workload_config(args, g):
    return {"workload": "LightGCN 3-layer d=64 BPR, one epoch = all %d mini-batches of 1024 + full-ranking top-20 eval of %d users x %d items, synthetic %s shape (%d users / %d items / %d train edges)"
                        % ((len(g.train_user) + 1023) // 1024, len(np.unique(g.test_user)), g.num_items, args.shape, g.num_users, g.num_items, len(g.train_user)),
            "batch_size": 1024, "layers": 3, "d": 64, "parallelism": "single GPU" if args.gpus == 1 else "row-partitioned x%d, per-layer all-gather; eval user-sharded" % args.gpus,
            "l2": "per-step working set (tables, gradients, Adam moments, CSR: > 400 MB) exceeds the 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------------------------
def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else any library prints there during the run
    (NCCL's version banner, torch warnings) has been diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    sys.stdout.flush()
    os.write(fd, data)


_REAL_STDOUT = None


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)            # stdout of this process (and of the native libraries it loads) -> stderr until emit()
    if args.impl == "reference":
        return run_reference(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from idgrec import _lib, datagen
    from utility.utility_data.data_loader import Data
    import utility.utility_function.tools as tools
    import utility.utility_train.batch_test as batch_test
    import utility.utility_train.trainer as trainer
    from models.LightGCN import LightGCN
    lib = _lib.lib()

    g = datagen.gen_graph(args.shape)
    cfg = dict(CFG)
    data = Data.from_arrays(g.num_users, g.num_items, g.train_user, g.train_item, g.test_user, g.test_item, cfg)
    tools.set_seed(2024)
    if world > 1:
        cfg["num_gpus"] = str(world)
    model = LightGCN(cfg, data, dev)
    model.to(dev)
    B = 1024
    ft = model.fused_trainer(1e-3, B)
    U, I, N, E, d, K = data.num_users, data.num_items, data.num_nodes, len(data.train_user), 64, 3
    nnz = model.Graph.csr.nnz

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def train_epoch(users, pos, neg):
        nb = 0
        for s in range(0, E, B):
            ft.step(users[s:s + B], pos[s:s + B], neg[s:s + B])
            nb += 1
            if args.batches and nb >= args.batches:
                break
        return nb

    def eval_once():
        return batch_test.Test(data, model, dev, cfg)

    # ---------------- value: samples resident on the device ----------------
    users, pos, neg = trainer.sample_epoch(data, dev)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        train_epoch(users, pos, neg)
        eval_once()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    l0, r0 = lib.idg_launch_count(), ft.replayed_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_train = t_eval = 0.0
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        ev[0].record()
        nb = train_epoch(users, pos, neg)
        ev[1].record()
        res = eval_once()
        ev[2].record()
        torch.cuda.synchronize()
        t_train += ev[0].elapsed_time(ev[1]) / 1e3
        t_eval += ev[1].elapsed_time(ev[2]) / 1e3
    barrier()
    wall = time.perf_counter() - w0
    launches = int(lib.idg_launch_count() - l0 + ft.replayed_launches - r0)
    clk = clocks.stop() if clocks else None
    t = torch.tensor([t_train + t_eval, t_train, t_eval, wall], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot, t_train, t_eval, wall = (float(x) / args.steps for x in t.tolist())

    # ---------------- e2e: through the public trainer API with host buffers ----------------
    h2d = 2 * E * 8                                           # negatives + permutation; the train edges are resident
    d2h = 4 * 4 + 3 * 2 * 8
    barrier()
    e0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 8))
    t_s = time.perf_counter()
    nxt = trainer.sample_epoch(data, dev)                    # host sampler + shuffle + pinned H2D (inside the timed region, not overlapped)
    t_first_sampling = time.perf_counter() - t_s
    phases = np.zeros(4)
    for i in range(n_e2e):
        u2, p2, n2 = nxt
        h0 = time.perf_counter()
        worker = trainer.EpochPrefetch(data) if i + 1 < n_e2e else None   # next epoch's host sampling on a worker thread, as in universal_trainer
        train_epoch(u2, p2, n2)                              # enqueues the epoch's steps
        h1 = time.perf_counter()
        if worker is not None:
            nxt = worker.finish(dev)                         # pinned H2D + device-side shuffle gather
        h2 = time.perf_counter()
        ft.pop_epoch_losses()                                # D2H of the epoch losses
        h3 = time.perf_counter()
        eval_once()                                          # D2H of the metric sums
        h4 = time.perf_counter()
        phases += [h1 - h0, h2 - h1, h3 - h2, h4 - h3]
    barrier()
    e2e = (time.perf_counter() - e0) / n_e2e
    if dist is not None:
        te = torch.tensor([e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = float(te.item())

    # ---------------- xl: the scaling target (BASELINE.json configs[4]) under the same launch ----------------
    xl = None
    if not args.no_xl and not args.batches:
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench_xl", os.path.join(REPO, "tools", "bench_xl.py"))
        bench_xl = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench_xl)
        barrier()
        try:
            xl = bench_xl.run(rank, world, dev, steps=args.xl_steps, warmup=5)
        except Exception as e:  # noqa: BLE001 -- a failed sub-record must not take the headline line with it
            xl = {"error": repr(e)[:400]}
        barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (one propagation layer) ----------------
    hbm, hbm_src = peaks()
    X, Y = ft.E0, torch.empty_like(ft.E0)
    n_s = 60
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_s)]
    for _ in range(5):
        model.Graph.spmm_layer(X, Y=Y)
    for a, b in evs:
        ft.m.mul_(1.0)  # touch other step tensors between launches like the real step does (L2 churn)
        a.record()
        model.Graph.spmm_layer(X, Y=Y)
        b.record()
    torch.cuda.synchronize()
    spmm_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    bytes_alg = 4 * (N + 1) + 8 * nnz + 8 * N * d
    bytes_gather = 4 * (N + 1) + 8 * nnz + 4 * nnz * d + 4 * N * d
    ach = bytes_alg / spmm_ms / 1e6
    roof = {"kernel": "spmm_kernel<16> (one propagation layer, d=64)", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            "traffic": None, "peak_source": hbm_src, "launch_ms": spmm_ms, "algorithmic_bytes": bytes_alg,
            "how": "CUDA events around %d eager launches on the launching stream, in this process after the timed epochs" % n_s,
            "gather_counted_GBs": bytes_gather / spmm_ms / 1e6,
            "note": "table (%.0f MB) is L2-resident at this shape: the kernel is bound by L2->SM gather bandwidth, not HBM; gather_counted_GBs is the L2-side rate" % (N * d * 4 / 1e6),
            "share_of_step": 6 * spmm_ms / (t_train * 1e3 / max(nb, 1))}
    try:
        tr = json.load(open(os.path.join(REPO, "profiles", "spmm_traffic.json")))
        roof["traffic"] = tr.get(args.shape)
    except Exception:
        pass

    line = {"metric": METRIC, "value": tot, "unit": "s/epoch", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, g), "clocks": clk,
            "e2e": {"value": e2e, "unit": "s/epoch", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "epochs_timed": n_e2e,
                    "first_epoch_host_sampling_s": t_first_sampling,
                    "host_wall_per_epoch_s": {"enqueue_steps": phases[0] / n_e2e, "sample_next_epoch_incl_wait_for_gpu": phases[1] / n_e2e,
                                              "read_losses": phases[2] / n_e2e, "Test": phases[3] / n_e2e},
                    "note": "host negative sampling + shuffle of epoch i+1 overlap the kernels of epoch i (as in universal_trainer); the first epoch's sampling is exposed and inside the timed region"},
            "gpu_launches": launches, "roofline": roof,
            "breakdown": {"train_s": t_train, "eval_s": t_eval, "train_batches": nb, "ms_per_train_batch": t_train * 1e3 / max(nb, 1),
                          "eval_users_per_s": len(data.test_dict) / t_eval, "wall_s_per_step": wall,
                          "recall@20": float(res["recall"][1]), "ndcg@20": float(res["ndcg"][1])}}
    line["scaling_note"] = ("value is a STRONG-scaling figure on a 37 MB table: the per-layer exchange (every GPU receives (G-1)/G x 36.9 MB) costs more "
                            "than the local rows save, so this shape does not shard; the scaling target is the xl sub-record (512 MB table)")
    if xl is not None:
        line["xl"] = xl
    if args.batches:
        line["invalid"] = "debug run with --batches %d" % args.batches
    if world == 1 and not args.batches:
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench_configs", os.path.join(REPO, "tools", "bench_configs.py"))
        bc = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bc)
        del ft, model
        torch.cuda.empty_cache()
        # the epoch WITHOUT the identical-result work skipping (last forward layer on batch rows only, first backward product over
        # batch columns only): all 6 propagation layers dense, as SURVEY 8(d) asks to be shown separately
        line["roofline"]["epoch_frac_of_hbm_peak_reference_dataflow"] = nb * (6 * bytes_alg + 7 * N * d * 4) / t_train / 1e9 / hbm

        def guarded(fn, *a, **k):
            try:
                return fn(*a, **k)
            except Exception as e:  # noqa: BLE001 -- a failed sub-record must not take the headline line with it
                return {"error": repr(e)[:400]}

        dense = guarded(bc.lightgcn_epoch_record, args.shape, dev, hbm, restrict_rows=False)
        if "error" not in dense:
            dense = {"epoch_s": dense["epoch_s"], "train_s": dense["train_s"], "ms_per_train_batch": dense["ms_per_train_batch"],
                     "epoch_frac_of_hbm_peak": dense["epoch_frac_of_hbm_peak"],
                     "note": "restrict_rows = 0: every step runs 6 dense propagation layers (the headline value skips work the loss never reads; results identical)"}
        line["unrestricted"] = dense
        if not args.no_configs:
            cfgs = {"LightGCN_yelp2018": guarded(bc.lightgcn_epoch_record, "yelp2018", dev, hbm)}
            for kind, shape in (("SimGCL", "yelp2018"), ("XSimGCL", "yelp2018"), ("NGCF", "amazon-book")):
                cfgs["%s_%s" % (kind, shape)] = guarded(bc.step_record, kind, shape, dev, hbm)
            line["configs"] = cfgs
        if not args.no_torch_baseline:
            from oracle import torch_cuda_baseline as T
            tb = guarded(T.epoch_estimate, g, dev)
            if "error" not in tb:
                tb["speedup_of_value"] = tb["value"] / tot
            line["torch_cuda_baseline"] = tb
    if world == 1 and not args.no_cpu_baseline:
        est = cpu_epoch_estimate(g, n_train_batches=3, n_eval_batches=2)
        line["cpu_baseline"] = {"value": est["epoch_s"], "unit": "s/epoch", "cores": est["cores"], "kind": "port", "sample": est["sample"],
                                "t_train_batch_s": est["t_train_batch_s"], "t_eval_batch_s": est["t_eval_batch_s"]}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
