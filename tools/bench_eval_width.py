"""Ranking time of one whole-test-set idg_eval_topk call at a given table width, on the amazon-book shape.

    python tools/bench_eval_width.py [--d 256] [--scale 0.1]           # tcgen05 candidate pass (d = 64 / 256)
    IDG_EVAL_IMPL=fma python tools/bench_eval_width.py --d 256         # the CUDA-core candidate pass, same inputs

Tables are Gaussian with heavy-tailed item norms (every 7th row x4), which is what a trained table looks like to the filter;
users = every user with a test row, masks = the training interactions.  Prints one JSON line.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "id-grec_b200"), REPO):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=256)
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--shape", default="amazon-book")
    ap.add_argument("--K", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    import scipy.sparse as sp
    from idgrec import datagen, ops
    g = datagen.gen_graph(a.shape)
    U, I = g.num_users, g.num_items
    net = sp.csr_matrix((np.ones(len(g.train_user), np.float32), (g.train_user, g.train_item)), shape=(U, I))
    net.sort_indices()
    mp = torch.from_numpy(net.indptr.astype(np.int32)).to(dev)
    mi = torch.from_numpy(net.indices.astype(np.int32)).to(dev)
    gen = torch.Generator(device=dev).manual_seed(5)
    Fu = torch.randn(U, a.d, generator=gen, device=dev) * a.scale
    Fi = torch.randn(I, a.d, generator=gen, device=dev) * a.scale
    Fi[::7] *= 4.0
    users = torch.from_numpy(np.unique(g.test_user).astype(np.int64)).to(dev)
    for _ in range(3):
        ids = ops.eval_topk(Fu, Fi, users, mp, mi, a.K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        ids = ops.eval_topk(Fu, Fi, users, mp, mi, a.K)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * len(users) * I * a.d
    print(json.dumps({"d": a.d, "impl": os.environ.get("IDG_EVAL_IMPL", "tc"), "users": int(len(users)), "items": I, "K": a.K, "scale": a.scale,
                      "ms_per_call": ms, "score_TFLOPs": flops / ms / 1e9, "ids_checksum": int(ids.sum().item())}))


if __name__ == "__main__":
    main()
