"""Workloads for the round-2 ncu captures (profiles/*_r2*): eager launches so the profiler sees every kernel.

    python tools/prof_r2.py lightgcn   # 3 eager LightGCN train steps, amazon-book shape (spmm_kernel variants, BPR kernels)
    python tools/prof_r2.py eval       # 400 graph-replayed train steps (so the table is not at its initial scale), then one Test()
    python tools/prof_r2.py infonce    # InfoNCE forward+backward at n = 1,923 rows (the SimGCL production size), tau = 0.2
    python tools/prof_r2.py simgcl     # 2 eager SimGCL steps on the yelp2018 shape, B = 2,048
    python tools/prof_r2.py ngcf       # 2 eager NGCF steps on the amazon-book shape
    python tools/prof_r2.py all        # everything above (except ngcf) once (launch list)
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "id-grec_b200"), REPO, os.path.join(REPO, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import bench_configs as bc  # noqa: E402


def lightgcn(dev, n_graph_steps=0, eager_steps=3, do_eval=False):
    import utility.utility_train.batch_test as batch_test
    import utility.utility_train.trainer as trainer
    cfg, g, data, model = bc._build("LightGCN", "amazon-book", dev)
    B = 1024
    users, pos, neg = trainer.sample_epoch(data, dev)
    if n_graph_steps:
        ft = model.fused_trainer(1e-3, B)
        for s in range(n_graph_steps):
            ft.step(users[s * B:(s + 1) * B], pos[s * B:(s + 1) * B], neg[s * B:(s + 1) * B])
    if eager_steps:
        cfg["cuda_graph"] = "0"
        model._fused = None
        ft = model.fused_trainer(1e-3, B)
        for s in range(eager_steps):
            ft.step(users[s * B:(s + 1) * B], pos[s * B:(s + 1) * B], neg[s * B:(s + 1) * B])
    if do_eval:
        batch_test.Test(data, model, dev, cfg)
    torch.cuda.synchronize()


def simgcl(dev):
    import utility.utility_train.trainer as trainer
    cfg, g, data, model = bc._build("SimGCL", "yelp2018", dev)
    cfg["cuda_graph"] = "0"
    B = 2048
    ft = model.fused_trainer(1e-3, B)
    users, pos, neg = trainer.sample_epoch(data, dev)
    for s in range(2):
        ft.step(users[s * B:(s + 1) * B], pos[s * B:(s + 1) * B], neg[s * B:(s + 1) * B])
    torch.cuda.synchronize()


def ngcf(dev):
    import utility.utility_train.trainer as trainer
    cfg, g, data, model = bc._build("NGCF", "amazon-book", dev)
    cfg["cuda_graph"] = "0"
    B = 1024
    ft = model.fused_trainer(float(cfg["learn_rate"]), B)
    users, pos, neg = trainer.sample_epoch(data, dev)
    for s in range(2):
        ft.step(users[s * B:(s + 1) * B], pos[s * B:(s + 1) * B], neg[s * B:(s + 1) * B])
    torch.cuda.synchronize()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    if what in ("lightgcn", "all"):
        lightgcn(dev)
    if what in ("eval", "all"):
        lightgcn(dev, n_graph_steps=400, eager_steps=0, do_eval=True)
    if what in ("infonce", "all"):
        print(bc.infonce_record(dev, 1923, 0.2))
    if what in ("simgcl", "all"):
        simgcl(dev)
    if what == "ngcf":
        ngcf(dev)


if __name__ == "__main__":
    main()
