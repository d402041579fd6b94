"""Experiment: does an L2 access-policy window on the gather table help the XL propagation layer (table 512 MB >> 126 MB L2)?

Times one full layer (1-GPU shape) and one layer over 1/8 of the rows (what a rank of 8 computes), each without a window and with
a persisting window of several hit ratios, and the 1/8 layer under concurrent write traffic (a 448 MB fill on a second stream:
what the seven peers push into this GPU's memory during an exchanged layer).

    python tools/exp_l2_window.py
"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "id-grec_b200"), REPO):
    sys.path.insert(0, p)


def main():
    from cuda.bindings import runtime as rt
    from idgrec import datagen
    from idgrec.graph import Graph, build_norm_adjacency
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    U = I = 1000000
    E, d = 100000000, 64
    N = U + I
    eu, ei = datagen.gen_edges_device(U, I, E, 2024, dev)
    csr = build_norm_adjacency(eu, ei, U, I, device=dev)
    del eu, ei
    full, part = Graph(csr), Graph(csr, 0, N // 8 // 128 * 128)
    X = torch.rand(N, d, device=dev)
    Y = torch.empty(N, d, device=dev)
    pollute = torch.empty(448 << 20, dtype=torch.uint8, device=dev)
    err, prop = rt.cudaGetDeviceProperties(0)
    out = {"persistingL2CacheMaxSize": int(prop.persistingL2CacheMaxSize), "accessPolicyMaxWindowSize": int(prop.accessPolicyMaxWindowSize),
           "l2CacheSize": int(prop.l2CacheSize)}
    st, side = torch.cuda.Stream(), torch.cuda.Stream()

    def window(ratio):
        attr = rt.cudaStreamAttrValue()
        if ratio is None:
            attr.accessPolicyWindow.num_bytes = 0
        else:
            rt.cudaDeviceSetLimit(rt.cudaLimit.cudaLimitPersistingL2CacheSize, int(prop.persistingL2CacheMaxSize))
            attr.accessPolicyWindow.base_ptr = X.data_ptr()
            attr.accessPolicyWindow.num_bytes = min(X.numel() * 4, int(prop.accessPolicyMaxWindowSize))
            attr.accessPolicyWindow.hitRatio = ratio
            attr.accessPolicyWindow.hitProp = rt.cudaAccessProperty.cudaAccessPropertyPersisting
            attr.accessPolicyWindow.missProp = rt.cudaAccessProperty.cudaAccessPropertyStreaming
        e = rt.cudaStreamSetAttribute(st.cuda_stream, rt.cudaStreamAttrID.cudaLaunchAttributeAccessPolicyWindow, attr)
        if ratio is None:
            rt.cudaCtxResetPersistingL2Cache()
        return str(e)

    def timed(g, with_pollution=False, reps=6):
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            if with_pollution:
                with torch.cuda.stream(side):
                    pollute.fill_(1)
            with torch.cuda.stream(st):
                a.record()
                g.spmm_layer(X, Y=Y)
                b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    for ratio in (None, 0.1, 0.18, 0.3, 1.0):
        rc = window(ratio)
        key = "none" if ratio is None else "hit%.2f" % ratio
        out[key] = {"rc": rc, "full_ms": timed(full), "eighth_ms": timed(part), "eighth_polluted_ms": timed(part, True)}
        print(key, out[key], flush=True)
    window(None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
