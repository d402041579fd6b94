"""Workload for the ncu capture of one propagation layer at the XL shape (1M x 1M / 100M edges, table 512 MB > L2):

    ncu --set full --clock-control none -k regex:spmm_kernel --launch-skip 3 -c 1 -o gpurun_out/spmm_xl python tools/prof_xl_spmm.py

Three warm launches, then the captured one; prints the live CUDA-event time of a plain layer for comparison (never a bench value
when run under the profiler).
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "id-grec_b200"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    from idgrec import datagen
    from idgrec.graph import Graph, build_norm_adjacency
    dev = torch.device("cuda:0")
    U = I = int(os.environ.get("XL_USERS", 1000000))
    E = int(os.environ.get("XL_EDGES", 100000000))
    eu, ei = datagen.gen_edges_device(U, I, E, 2024, dev)
    csr = build_norm_adjacency(eu, ei, U, I, device=dev)
    del eu, ei
    g = Graph(csr)
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    X = (torch.rand(U + I, 64, generator=gen, device=dev) * 2 - 1) * 0.05
    Y = torch.empty_like(X)
    n = int(os.environ.get("XL_LAUNCHES", 4))
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.spmm_layer(X, Y=Y)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    print("layer ms:", [round(a.elapsed_time(b), 3) for a, b in evs], "nnz", csr.nnz)


if __name__ == "__main__":
    main()
