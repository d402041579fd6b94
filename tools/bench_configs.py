"""Driver-visible numbers for BASELINE.json configs 1 / 3 / 4 (bench.py folds them into its one JSON line as
``configs``): LightGCN epoch on the yelp2018 shape, SimGCL / XSimGCL fused steps on the yelp2018 shape (B = 2,048,
shipped hyper-parameters), NGCF fused step on the amazon-book shape -- each through the model classes' public
``fused_trainer`` (the call universal_trainer makes), timed with CUDA events over real sampled batches, with the
algorithmic bytes of SURVEY.md section 8(d)'s reference dataflow and the fraction of the measured HBM peak they amount to.
"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "id-grec_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

BASE = {"embedding_size": "64", "test_batch_size": "1024", "reg_lambda": "0.0001", "GCN_layer": "3", "top_K": "[10, 20]",
        "sparsity_test": "0", "dataset": "synthetic", "interval": "1", "training_epochs": "1", "early_stopping": "10"}
MODEL_CFG = {
    "LightGCN": {"batch_size": "1024", "learn_rate": "0.001"},
    "SimGCL": {"batch_size": "2048", "learn_rate": "0.001", "ssl_lambda": "0.5", "temperature": "0.2", "epsilon": "0.05", "top_K": "[20, 40]"},
    "XSimGCL": {"batch_size": "2048", "learn_rate": "0.001", "ssl_lambda": "0.2", "temperature": "0.15", "epsilon": "0.2", "cl_layer": "1"},
    "NGCF": {"batch_size": "1024", "learn_rate": "0.0001", "mess_dropout": "True", "mess_drop_prob": "[0.1, 0.1, 0.1]", "node_dropout": "False",
             "node_drop_prob": "0.1", "layer_size": "[64, 64, 64]"},
}


def spmm_bytes(N, nnz, d=64):
    """one propagation layer = indptr + (col, val) + read X once + write Y once."""
    return 4 * (N + 1) + 8 * nnz + 8 * N * d


def _events(n):
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


def _build(kind, shape, dev):
    import importlib
    from idgrec import datagen
    from utility.utility_data.data_loader import Data
    import utility.utility_function.tools as tools
    cfg = dict(BASE)
    cfg.update(MODEL_CFG[kind])
    g = datagen.gen_graph(shape)
    data = Data.from_arrays(g.num_users, g.num_items, g.train_user, g.train_item, g.test_user, g.test_item, cfg)
    tools.set_seed(2024)
    model = getattr(importlib.import_module("models." + kind), kind)(cfg, data, dev)
    model.to(dev)
    return cfg, g, data, model


def step_record(kind, shape, dev, hbm, steps=120, warmup=40):
    """ms per fused train step of `kind` on a synthetic graph of `shape`, batches from the reference's own sampler."""
    import utility.utility_train.trainer as trainer
    cfg, g, data, model = _build(kind, shape, dev)
    B = int(cfg["batch_size"])
    ft = model.fused_trainer(float(cfg["learn_rate"]), B)
    users, pos, neg = trainer.sample_epoch(data, dev)
    nb = min(steps + warmup, len(users) // B)
    steps = nb - warmup
    for s in range(warmup):
        ft.step(users[s * B:(s + 1) * B], pos[s * B:(s + 1) * B], neg[s * B:(s + 1) * B])
    torch.cuda.synchronize()
    a, b = _events(2)
    a.record()
    for s in range(warmup, nb):
        ft.step(users[s * B:(s + 1) * B], pos[s * B:(s + 1) * B], neg[s * B:(s + 1) * B])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    N, nnz, d = data.num_nodes, model.Graph.csr.nnz, 64
    sp = spmm_bytes(N, nnz, d)
    adam = 7 * N * d * 4
    if kind == "SimGCL":      # 3 propagations fwd + 3 bwd = 18 SpMM, 6 noise passes (read noise, rewrite layer), Adam
        ref_bytes = 18 * sp + 6 * 2 * N * d * 4 + adam
        note = "reference dataflow: 18 SpMM + 6 noise passes + Adam; here the three backward chains share one (9 -> 3 SpMM) and the last forward layers run on batch rows only"
    elif kind == "XSimGCL":   # 1 propagation fwd + bwd = 6 SpMM, 3 noise passes
        ref_bytes = 6 * sp + 3 * 2 * N * d * 4 + adam
        note = "reference dataflow: 6 SpMM + 3 noise passes + Adam"
    elif kind == "NGCF":      # per layer fwd: SpMM + 6 elementwise N x 64 passes; bwd the same again; Adam
        ref_bytes = 6 * sp + 2 * 3 * 6 * 2 * N * d * 4 + adam
        note = "reference dataflow: 6 SpMM (graph with self loops) + 6 elementwise [N,64] passes per layer each way + Adam; dense 2 x 2 N 64 64 flop per layer fwd"
    else:
        ref_bytes = 6 * sp + adam
        note = "reference dataflow: 6 SpMM + Adam"
    rec = {"model": kind, "shape": shape, "batch": B, "ms_per_train_step": ms, "steps_timed": steps,
           "algorithmic_bytes_per_step_reference_dataflow": ref_bytes, "achieved_GBs": ref_bytes / ms / 1e6,
           "frac_of_hbm_peak": ref_bytes / ms / 1e6 / hbm, "nnz": nnz, "nodes": N, "note": note,
           "loss_sums_over_all_steps": [float(x) for x in ft.pop_epoch_losses()]}
    if kind == "NGCF":
        rec["dense_flops_per_step"] = 3 * 3 * 2 * 2 * N * 64 * 64      # fwd + 2x bwd, 3 layers, two 64x64 products
    if kind in ("SimGCL", "XSimGCL"):
        rec["infonce"] = infonce_record(dev, int(0.94 * B), float(cfg["temperature"]))
    del ft, model, data
    torch.cuda.empty_cache()
    return rec


def infonce_record(dev, n, tau, d=64):
    """One InfoNCE forward+backward call at the production row count (~1,923 unique users of a 2,048 batch)."""
    from idgrec import _lib
    l = _lib.lib()
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    N = n + 64
    V1 = torch.randn(N, d, generator=gen, device=dev) * 0.1
    V2 = V1 + 0.03 * torch.randn(N, d, generator=gen, device=dev)
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    g1, g2 = torch.zeros_like(V1), torch.zeros_like(V2)
    ws = torch.empty(int(l.idg_infonce_workspace_bytes(n, d)), dtype=torch.uint8, device=dev)
    loss = torch.zeros(1, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def call():
        _lib.check(l.idg_infonce_fwd_bwd(V1.data_ptr(), V2.data_ptr(), idx.data_ptr(), n, d, tau, 0.5, loss.data_ptr(), g1.data_ptr(), g2.data_ptr(),
                                         ws.data_ptr(), st), "idg_infonce_fwd_bwd")
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    evs = [_events(2) for _ in range(30)]
    for a, b in evs:
        a.record()
        call()
        b.record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
    flops = 4 * 2.0 * n * n * d          # two flash passes, each the score product and the E V product: 4 n x n x d products
    rec = {"rows": n, "ms_per_call": ms, "flops_fp32_equivalent": flops, "TFLOPs_fp32_equivalent": flops / ms / 1e9,
           "passes": "3xTF32 split products on tcgen05 (x3 tensor-core work per fp32-equivalent flop); E = exp(S/tau) stays in TMEM (csrc/infonce_flash.cu)"}
    try:
        prof = json.load(open(os.path.join(REPO, "profiles", "infonce_flash_n1923_r2_ncu.json")))
        rec["tensor_pipe_active_pct_ncu"] = prof.get("tensor_pipe_active_pct")
        rec["kernel_us_ncu"] = prof.get("kernel_us")
        rec["ncu_source"] = "profiles/infonce_flash_n1923_r2_ncu.json"
    except Exception:
        pass
    return rec


def lightgcn_epoch_record(shape, dev, hbm, restrict_rows=True):
    """One LightGCN epoch (all mini-batches + full-ranking eval) on `shape`, CUDA events; the layer roofline of that shape."""
    import utility.utility_train.batch_test as batch_test
    import utility.utility_train.trainer as trainer
    cfg, g, data, model = _build("LightGCN", shape, dev)
    if not restrict_rows:
        cfg["restrict_rows"] = "0"
    B = 1024
    ft = model.fused_trainer(1e-3, B)
    users, pos, neg = trainer.sample_epoch(data, dev)
    E = len(users)

    def epoch():
        for s in range(0, E, B):
            ft.step(users[s:s + B], pos[s:s + B], neg[s:s + B])

    epoch()                              # untimed: captures the step graph and lets the clocks reach their loaded state (an idle
    batch_test.Test(data, model, dev, cfg)   # B200 sits at 120 MHz; a 50-step warm-up measured 1.5x slower epochs when run first)
    torch.cuda.synchronize()
    a, b, c = _events(3)
    a.record()
    epoch()
    b.record()
    res = batch_test.Test(data, model, dev, cfg)
    c.record()
    torch.cuda.synchronize()
    t_train, t_eval = a.elapsed_time(b) / 1e3, b.elapsed_time(c) / 1e3
    N, nnz, d = data.num_nodes, model.Graph.csr.nnz, 64
    X, Y = ft.E0, torch.empty_like(ft.E0)
    evs = [_events(2) for _ in range(40)]
    for _ in range(5):
        model.Graph.spmm_layer(X, Y=Y)
    for x, y in evs:
        ft.m.mul_(1.0)
        x.record()
        model.Graph.spmm_layer(X, Y=Y)
        y.record()
    torch.cuda.synchronize()
    layer_ms = float(np.mean([x.elapsed_time(y) for x, y in evs]))
    sp = spmm_bytes(N, nnz, d)
    nb = (E + B - 1) // B
    rec = {"model": "LightGCN", "shape": shape, "restrict_rows": bool(restrict_rows), "epoch_s": t_train + t_eval, "train_s": t_train, "eval_s": t_eval,
           "train_batches": nb, "ms_per_train_batch": t_train * 1e3 / nb, "eval_users_per_s": len(data.test_dict) / t_eval,
           "layer_ms": layer_ms, "layer_alg_bytes": sp, "layer_GBs": sp / layer_ms / 1e6, "layer_frac_of_hbm_peak": sp / layer_ms / 1e6 / hbm,
           "epoch_alg_bytes_reference_dataflow": nb * (6 * sp + 7 * N * d * 4),
           "epoch_frac_of_hbm_peak": nb * (6 * sp + 7 * N * d * 4) / t_train / 1e9 / hbm,
           "recall@20": float(res["recall"][1]), "ndcg@20": float(res["ndcg"][1])}
    del ft, model, data
    torch.cuda.empty_cache()
    return rec


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    try:
        hbm = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    out = {"yelp2018_lightgcn": lightgcn_epoch_record("yelp2018", dev, hbm)}
    for kind, shape in (("SimGCL", "yelp2018"), ("XSimGCL", "yelp2018"), ("NGCF", "amazon-book")):
        out["%s_%s" % (kind, shape)] = step_record(kind, shape, dev, hbm)
    print(json.dumps(out, indent=1))
