"""Per-kernel timings on synthetic graphs of the reference shapes (CUDA events, L2 flushed
between timed launches).  Development aid; the judged numbers come from bench.py."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "id-grec_b200"))
from idgrec import datagen, ops, _lib  # noqa: E402
from idgrec.graph import Graph, build_norm_adjacency  # noqa: E402


def timed(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="amazon-book")
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    g = datagen.gen_graph(args.shape)
    U, I = g.num_users, g.num_items
    N, d, K = U + I, 64, 3
    t0 = time.time()
    csr = build_norm_adjacency(g.train_user, g.train_item, U, I, device=dev)
    torch.cuda.synchronize()
    t_csr = time.time() - t0
    G = Graph(csr)
    nnz = csr.nnz
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    X = (torch.rand(N, d, device=dev) - 0.5) * 0.1
    Y, acc = torch.empty_like(X), torch.empty_like(X)
    out = {"shape": args.shape, "U": U, "I": I, "nnz": nnz, "csr_build_s": t_csr}
    bytes_alg = 4 * (N + 1) + 8 * nnz + 8 * N * d
    bytes_gather = 4 * (N + 1) + 8 * nnz + 4 * nnz * d + 4 * N * d
    for name, fl in (("spmm_flushed", flush), ("spmm_warm", None)):
        med, mn = timed(lambda: G.spmm_layer(X, Y=Y), flush=fl)
        out[name] = {"ms": med, "min_ms": mn, "alg_GBs": bytes_alg / med / 1e6, "gather_GBs": bytes_gather / med / 1e6,
                     "frac_hbm_alg": bytes_alg / med / 1e6 / hbm}
    med, _ = timed(lambda: G.propagate_fwd(X, K, True, out_mean=acc), flush=None)
    out["propagate_fwd_ms"] = med
    Gd = torch.randn(N, d, device=dev) * 1e-3
    med, _ = timed(lambda: G.propagate_bwd(Gd, K, True, out=acc), flush=None)
    out["propagate_bwd_ms"] = med
    # BPR
    B = 1024
    l = _lib.lib()
    users = torch.randint(0, U, (B,), device=dev)
    pos = torch.randint(0, I, (B,), device=dev)
    neg = torch.randint(0, I, (B,), device=dev)
    ws = torch.empty(int(l.idg_bpr_workspace_bytes(B)), dtype=torch.uint8, device=dev)
    loss = torch.empty(2, device=dev)
    Gz = torch.zeros(N, d, device=dev)
    gE = torch.zeros(N, d, device=dev)
    st = lambda: torch.cuda.current_stream().cuda_stream

    def bpr():
        _lib.check(l.idg_bpr_forward(X.data_ptr(), X.data_ptr(), users.data_ptr(), pos.data_ptr(), neg.data_ptr(), B, U, N, d, 1e-4, 7, loss.data_ptr(), ws.data_ptr(), st()))
        _lib.check(l.idg_bpr_backward(X.data_ptr(), B, d, 7, None, Gz.data_ptr(), 0.0, None, ws.data_ptr(), st()))
        _lib.check(l.idg_bpr_finish(X.data_ptr(), gE.data_ptr(), Gz.data_ptr(), B, d, 1e-4, None, None, ws.data_ptr(), st()))
    out["bpr_fwd_bwd_finish_ms"], _ = timed(bpr)
    m, v = torch.zeros_like(X), torch.zeros_like(X)
    out["adam_ms"], _ = timed(lambda: ops.adam_step(X, Gd, m, v, 1e-3, 1))
    # InfoNCE alone (n ~ unique users of a 2048 batch) and the contrastive models' fused steps
    n = 1900
    idx = torch.sort(torch.randperm(U, device=dev)[:n]).values
    V1, V2 = torch.randn(N, d, device=dev), torch.randn(N, d, device=dev)
    g1 = torch.zeros(N, d, device=dev)
    wsn = torch.empty(int(l.idg_infonce_workspace_bytes(n, d)), dtype=torch.uint8, device=dev)
    lossn = torch.zeros(1, device=dev)
    out["infonce_n1900_ms"], _ = timed(lambda: _lib.check(l.idg_infonce_fwd_bwd(V1.data_ptr(), V2.data_ptr(), idx.data_ptr(), n, d, 0.2, 0.5, lossn.data_ptr(), g1.data_ptr(), g1.data_ptr(), wsn.data_ptr(), st())))
    from idgrec.engine import FusedTrainer
    Bc = 2048
    uc, pc, nc = (torch.randint(0, U, (Bc,), device=dev), torch.randint(0, I, (Bc,), device=dev), torch.randint(0, I, (Bc,), device=dev))
    for kind, kw in (("LightGCN", {}), ("SimGCL", dict(ssl_lambda=0.5, temperature=0.2, eps=0.05)), ("XSimGCL", dict(ssl_lambda=0.2, temperature=0.15, eps=0.2, cl_layer=1))):
        ftc = FusedTrainer(kind, G, X.clone(), U, 3, 1e-4, 1e-3, max_batch=Bc, use_cuda_graph=False, **kw)
        out["step_%s_B2048_eager_ms" % kind], _ = timed(lambda: ftc.step(uc, pc, nc), iters=10, warm=2)
        del ftc
        ftg = FusedTrainer(kind, G, X.clone(), U, 3, 1e-4, 1e-3, max_batch=Bc, use_cuda_graph=True, **kw)
        out["step_%s_B2048_graph_ms" % kind], _ = timed(lambda: ftg.step(uc, pc, nc), iters=20, warm=3)
        del ftg
    # NGCF: one autograd step (3 x (SpMM + fused dense kernel) fwd/bwd, 256-d BPR) on the with-self graph
    if os.environ.get("IDG_BENCH_NGCF", "1") == "1":
        csr_s = build_norm_adjacency(g.train_user, g.train_item, U, I, add_self=True, device=dev)
        Gs = Graph(csr_s)
        Eng = X.clone().requires_grad_(True)
        Ws = [[torch.nn.init.xavier_uniform_(torch.empty(64, 64, device=dev)).requires_grad_(True), torch.zeros(1, 64, device=dev, requires_grad=True),
               torch.nn.init.xavier_uniform_(torch.empty(64, 64, device=dev)).requires_grad_(True), torch.zeros(1, 64, device=dev, requires_grad=True)] for _ in range(3)]
        ub, pb, nb_ = torch.randint(0, U, (1024,), device=dev), torch.randint(0, I, (1024,), device=dev), torch.randint(0, I, (1024,), device=dev)

        def ngcf_step():
            ego, outs = Eng, [Eng]
            for Wg, bg, Wb, bb in Ws:
                keep = (torch.rand_like(ego.detach()) >= 0.1).float()
                ego, o = ops.ngcf_layer(ego, Wg, bg, Wb, bb, Gs, keep, 0.1)
                outs.append(o)
            fin = torch.cat(outs, dim=1)
            loss = ops.bpr_reg_loss(fin, fin, ub, pb, nb_, U, 0.0, 0)[0]
            loss.backward()
        out["step_NGCF_B1024_autograd_ms"], _ = timed(ngcf_step, iters=10, warm=2)
        del Gs, csr_s
    # eval
    import scipy.sparse as sp
    net = sp.csr_matrix((np.ones(len(g.train_user)), (g.train_user, g.train_item)), shape=(U, I))
    net.sort_indices()
    mp = torch.from_numpy(net.indptr.astype(np.int32)).to(dev)
    mi = torch.from_numpy(net.indices.astype(np.int32)).to(dev)
    F = torch.randn(N, d, device=dev) * 0.4
    tu = torch.arange(U, device=dev)
    wse = torch.empty(int(l.idg_eval_workspace_bytes(U, I, d, 20)), dtype=torch.uint8, device=dev)
    med, _ = timed(lambda: ops.eval_topk(F[:U], F[U:], tu, mp, mi, 20, ws=wse), iters=3, warm=1)
    out["eval_topk_ms"] = med
    out["eval_users_per_s"] = U / med * 1e3
    out["eval_TFLOPs"] = 2.0 * U * I * d / med / 1e9
    flag = wse[256:260].view(torch.int32).item()
    out["eval_flagged_users"] = int(flag)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
