"""Timings of the SURVEY section-8(f) rows on one B200 (CUDA events; development aid, not the judged bench):
batch x batch loss kernels at the configured batch sizes, the autograd train step of the models built on them,
and the host legs of one epoch (text parse, exact negative sampler + shuffle + H2D)."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "id-grec_b200"))
from idgrec import datagen, ops  # noqa: E402


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    import importlib
    import utility.utility_function.tools as tools
    import utility.utility_train.trainer as trainer
    from utility.utility_data.data_loader import Data
    dev = torch.device("cuda:0")
    out = {}
    # ---- loss kernels, forward + backward in one call
    for n in (2048, 4096):
        X = torch.randn(n, 64, device=dev).requires_grad_(True)
        Y = (torch.randn(n, 64, device=dev) + 0.5 * X.detach()).requires_grad_(True)
        for kind, p0, p1 in (("lightccf", 0.22, 0.0), ("lightcscf", 0.2, 0.7), ("sccf_down", 0.1, float(n * n) / 2), ("sccf_up", 0.1, 0.0),
                             ("align", 0.0, 0.0), ("uniform", 0.0, 0.0)):
            out["pair_%s_n%d_ms" % (kind, n)] = timed(lambda: ops.pair_loss(kind, X, None if kind == "uniform" else Y, p0, p1))
        idx = torch.randint(0, 50000, (n,), device=dev)
        T = torch.randn(50000, 64, device=dev)
        G = torch.randn(n, 64, device=dev)
        gT = torch.zeros_like(T)
        from idgrec import _lib
        out["scatter_add_n%d_ms" % n] = timed(lambda: _lib.check(_lib.lib().idg_scatter_add_rows(_lib.ptr(G), _lib.ptr(idx), n, 64, _lib.ptr(gT), _lib.cur_stream())))
    # ---- one autograd train step of the added models
    graphs = {}
    for shape, kind, over in (("amazon-book", "LightCCF", {}), ("amazon-book", "LightCSCF", {}), ("amazon-book", "NGCF", {}),
                              ("yelp2018", "DirectAU", {}), ("yelp2018", "SCCF", {"encoder": "LightGCN"}), ("yelp2018", "SGL", {}), ("yelp2018", "EGCF", {})):
        g = graphs.get(shape) or graphs.setdefault(shape, datagen.gen_graph(shape))
        cfg = tools.read_configuration(os.path.join(REPO, "id-grec_b200", "configure", kind + ".txt"), kind)
        cfg.update(over)
        data = Data.from_arrays(g.num_users, g.num_items, g.train_user, g.train_item, g.test_user, g.test_item, cfg)
        tools.set_seed(2024)
        m = getattr(importlib.import_module("models." + kind), kind)(cfg, data, dev)
        m.to(dev)
        opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        B = int(cfg["batch_size"])
        rng = np.random.default_rng(0)
        e = rng.integers(0, len(g.train_user), B)
        bu, bp = (torch.from_numpy(a[e]).to(dev) for a in (g.train_user, g.train_item))
        bn = torch.from_numpy(rng.integers(0, g.num_items, B)).to(dev)
        extra = ()
        if kind == "SGL":
            extra = tuple(tools.convert_sp_mat_to_sp_tensor(tools.create_adj_mat(data.user_item_net, "ed", 0.1)).to(dev) for _ in range(2))

        def step():
            ll = m(bu, bp, bn, *extra)
            opt.zero_grad()
            torch.stack([l.reshape(()) for l in ll]).sum().backward()
            opt.step()
        out["step_%s_%s_B%d_ms" % (kind, shape, B)] = timed(step, iters=10, warm=3)
        ft = m.fused_trainer(1e-3, B) if callable(getattr(m, "fused_trainer", None)) else None
        if ft is not None:
            out["step_fused_%s_%s_B%d_ms" % (kind, shape, B)] = timed(lambda: ft.step(bu, bp, bn), iters=20, warm=3)
        elif getattr(m, "graph_capturable", False):
            from idgrec.graphed import GraphedStep
            gs = GraphedStep(m, 1e-3, B)
            out["step_graph_%s_%s_B%d_ms" % (kind, shape, B)] = timed(lambda: gs.step(bu, bp, bn), iters=20, warm=3)
            del gs
        if kind == "LightCCF":
            # host legs of one epoch at the amazon-book shape
            t0 = time.perf_counter()
            trainer.sample_epoch(data, dev)
            out["sample_epoch_first_s"] = time.perf_counter() - t0
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                trainer.sample_epoch(data, dev)
                ts.append(time.perf_counter() - t0)
            out["sample_epoch_s"] = float(np.median(ts))
            with tempfile.TemporaryDirectory() as td:
                datagen.write_dataset(td, "ab", g)
                t0 = time.perf_counter()
                r = ops.parse_ratings(os.path.join(td, "ab", "train.txt"))
                out["parse_train_txt_s"] = time.perf_counter() - t0
                out["parse_train_txt_pairs"] = int(len(r[2]))
        del m, opt, data
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(REPO, "gpurun_out", "bench_next.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
