"""One propagation layer timed at the yelp2018 / amazon-book shapes (CUDA events, other step tensors touched between launches)."""
import json, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "id-grec_b200"))
from idgrec import datagen
from idgrec.graph import Graph, build_norm_adjacency
dev = torch.device("cuda:0")
out = {}
for shape in ("yelp2018", "amazon-book"):
    g = datagen.gen_graph(shape)
    U, I = g.num_users, g.num_items
    csr = build_norm_adjacency(g.train_user, g.train_item, U, I, device=dev)
    G = Graph(csr)
    X = (torch.rand(U + I, 64, device=dev) - 0.5) * 0.1
    Y = torch.empty_like(X)
    churn = torch.zeros(64 << 20, device=dev)
    for _ in range(200):
        G.spmm_layer(X, Y=Y)
    evs = []
    for _ in range(60):
        churn.mul_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); G.spmm_layer(X, Y=Y); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    out[shape] = round(float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3, 2)
print(os.environ.get("IDG_SPMM_WARPS", "8"), json.dumps(out))
