"""idg_ngcf_dense_bwd against an fp64 torch restatement of the same algebra (csrc/ngcf.cu header), output by output.

    python tools/diag_ngcf_bwd.py [N]          # IDG_NGCF_BWD=fma for the CUDA-core kernel
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "id-grec_b200"), REPO):
    sys.path.insert(0, p)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    dev = torch.device("cuda:0")
    from idgrec import _lib
    from idgrec._lib import check, ptr, cur_stream
    l = _lib.lib()
    g = torch.Generator(device=dev).manual_seed(3)
    rn = lambda *s: torch.randn(*s, generator=g, device=dev)
    E, side = rn(N, 64) * 0.3, rn(N, 64) * 0.3
    Wg, Wb = rn(64, 64) * 0.2, rn(64, 64) * 0.2
    bg, bb = rn(64) * 0.1, rn(64) * 0.1
    keep = (torch.rand(N, 64, generator=g, device=dev) < 0.9).float()
    p = 0.1
    dO_full = rn(N, 256)
    dO = dO_full[:, 64:128]
    dDx = rn(N, 64) * 0.5
    S = torch.empty(N, 64, device=dev); D = torch.empty(N, 64, device=dev); out = torch.empty(N, 256, device=dev)
    check(l.idg_ngcf_dense_fwd(ptr(E), ptr(side), ptr(Wg), ptr(bg), ptr(Wb), ptr(bb), ptr(keep), p, N, ptr(S), ptr(D), out[:, 64:].data_ptr(), 256, cur_stream()), "fwd")
    dside = torch.empty(N, 64, device=dev); dEd = torch.empty(N, 64, device=dev)
    dWg = torch.empty(64, 64, device=dev); dWb = torch.empty(64, 64, device=dev); db = torch.empty(64, device=dev)
    ws = torch.empty(int(l.idg_ngcf_workspace_bytes()), dtype=torch.uint8, device=dev)
    check(l.idg_ngcf_dense_bwd(ptr(E), ptr(side), ptr(Wg), ptr(Wb), ptr(keep), p, ptr(S), ptr(D), dO.data_ptr(), 256, ptr(dDx), N, ptr(dside), ptr(dEd),
                               ptr(dWg), ptr(dWb), ptr(db), ptr(ws), cur_stream()), "bwd")
    torch.cuda.synchronize()
    f = lambda t: t.double()
    Dd, Sd = f(D), f(S)
    nrm = Dd.norm(dim=1, keepdim=True).clamp_min(1e-12)
    O = Dd / nrm
    dD = f(dDx) + (f(dO) - O * (O * f(dO)).sum(1, keepdim=True)) / nrm
    dS = dD * f(keep) / (1 - p) * torch.where(Sd > 0, 1.0, 0.2)
    Wcat = torch.cat([f(Wg), f(Wb)], 0)
    Z = torch.cat([f(side), f(E) * f(side)], 1)
    dZ = dS @ Wcat.T
    ref = {"dside": dZ[:, :64] + dZ[:, 64:] * f(E), "dE_direct": dZ[:, 64:] * f(side), "dWg": (Z.T @ dS)[:64], "dWb": (Z.T @ dS)[64:], "db": dS.sum(0)}
    got = {"dside": dside, "dE_direct": dEd, "dWg": dWg, "dWb": dWb, "db": db}
    for k in ref:
        err = (f(got[k]) - ref[k]).abs().max().item()
        print("%-10s max|err| %.3e  ref scale %.3e  rel %.3e" % (k, err, ref[k].abs().max().item(), err / ref[k].abs().max().item()))
    bad = (f(dside) - ref["dside"]).abs()
    rows = bad.max(1).values
    print("dside: rows with err > 1e-4:", int((rows > 1e-4).sum()), "first", torch.nonzero(rows > 1e-4)[:8].flatten().tolist())
    cols = bad.max(0).values
    print("dside: cols with err > 1e-4:", torch.nonzero(cols > 1e-4).flatten().tolist()[:64])
    wb = (f(torch.cat([dWg, dWb], 0)) - torch.cat([ref["dWg"], ref["dWb"]], 0)).abs()
    print("dW: rows bad", torch.nonzero(wb.max(1).values > 1e-3).flatten().tolist()[:128])
    print("dW: cols bad", torch.nonzero(wb.max(0).values > 1e-3).flatten().tolist()[:64])


if __name__ == "__main__":
    main()
