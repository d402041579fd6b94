"""Two eager train steps of one model at a reference shape (for `ncu --metrics gpu__time_duration.sum` launch lists).
    python tools/prof_step.py NGCF amazon-book"""
import importlib
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "id-grec_b200"))
from idgrec import datagen  # noqa: E402
import utility.utility_function.tools as tools  # noqa: E402
from utility.utility_data.data_loader import Data  # noqa: E402

kind, shape = sys.argv[1], sys.argv[2]
dev = torch.device("cuda:0")
g = datagen.gen_graph(shape)
cfg = tools.read_configuration(os.path.join(REPO, "id-grec_b200", "configure", kind + ".txt"), kind)
cfg["cuda_graph"] = "0"          # eager launches so that the profiler sees every kernel
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
ft = m.fused_trainer(1e-3, B) if callable(getattr(m, "fused_trainer", None)) else None
torch.cuda.synchronize()
if ft is not None:
    ft.step(bu, bp, bn)            # allocations, first-use attribute calls
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("steps")
    for _ in range(2):
        ft.step(bu, bp, bn)
else:
    torch.cuda.nvtx.range_push("steps")
    for _ in range(2):
        ll = m(bu, bp, bn)
        opt.zero_grad()
        torch.stack([l.reshape(()) for l in ll]).sum().backward()
        opt.step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
