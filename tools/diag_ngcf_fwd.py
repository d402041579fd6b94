"""idg_ngcf_dense_fwd against an fp64 torch restatement (models/NGCF.py:87-106), output by output, plus CUDA-event timings of the
forward and backward dense kernels at the given row count.

    python tools/diag_ngcf_fwd.py [N]
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
    g = torch.Generator(device=dev).manual_seed(5)
    rn = lambda *s: torch.randn(*s, generator=g, device=dev)
    E, side = rn(N, 64) * 0.3, rn(N, 64) * 0.3
    Wg, Wb = rn(64, 64) * 0.2, rn(64, 64) * 0.2
    bg, bb = rn(64) * 0.1, rn(64) * 0.1
    keep = (torch.rand(N, 64, generator=g, device=dev) < 0.9).float()
    p = 0.1
    S = torch.empty(N, 64, device=dev); D = torch.empty(N, 64, device=dev); out = torch.zeros(N, 256, device=dev)
    fwd = lambda: check(l.idg_ngcf_dense_fwd(ptr(E), ptr(side), ptr(Wg), ptr(bg), ptr(Wb), ptr(bb), ptr(keep), p, N, ptr(S), ptr(D), out[:, 64:].data_ptr(), 256,
                                             cur_stream()), "fwd")
    fwd()
    torch.cuda.synchronize()
    f = lambda t: t.double()
    Sd = f(side) @ f(Wg) + f(bg) + (f(E) * f(side)) @ f(Wb) + f(bb)
    Dd = torch.where(Sd > 0, Sd, 0.2 * Sd) * f(keep) / (1 - p)
    Od = Dd / Dd.norm(dim=1, keepdim=True).clamp_min(1e-12)
    for name, got, ref in (("S", S, Sd), ("D", D, Dd), ("O", out[:, 64:128], Od)):
        err = (f(got) - ref).abs().max().item()
        print("%-3s max|err| %.3e  ref scale %.3e  rel %.3e" % (name, err, ref.abs().max().item(), err / ref.abs().max().item()))
    assert float(out[:, :64].abs().max()) == 0 and float(out[:, 128:].abs().max()) == 0, "wrote outside its 64-column block"
    # timings
    dO = rn(N, 256); dDx = rn(N, 64)
    dside = torch.empty(N, 64, device=dev); dEd = torch.empty(N, 64, device=dev)
    dWg = torch.empty(64, 64, device=dev); dWb = torch.empty(64, 64, device=dev); db = torch.empty(64, device=dev)
    ws = torch.empty(int(l.idg_ngcf_workspace_bytes()), dtype=torch.uint8, device=dev)
    bwd = lambda: check(l.idg_ngcf_dense_bwd(ptr(E), ptr(side), ptr(Wg), ptr(Wb), ptr(keep), p, ptr(S), ptr(D), dO[:, 64:].data_ptr(), 256, ptr(dDx), N, ptr(dside),
                                             ptr(dEd), ptr(dWg), ptr(dWb), ptr(db), ptr(ws), cur_stream()), "bwd")
    import ctypes as C
    bits = torch.zeros(N, 2, dtype=torch.int32, device=dev)
    check(l.idg_ngcf_keep_bits(ptr(bits), N, 1, (C.c_float * 1)(0.9), 5, None, cur_stream()), "keep_bits")
    fwd_b = lambda: check(l.idg_ngcf_dense_fwd_bits(ptr(E), ptr(side), ptr(Wg), ptr(bg), ptr(Wb), ptr(bb), ptr(bits), p, N, None, ptr(D), out[:, 64:].data_ptr(), 256,
                                                    cur_stream()), "fwd_bits")
    bwd_b = lambda: check(l.idg_ngcf_dense_bwd_bits(ptr(E), ptr(side), ptr(Wg), ptr(Wb), ptr(bits), p, ptr(D), dO[:, 64:].data_ptr(), 256, ptr(dDx), N, ptr(dside),
                                                    ptr(dEd), ptr(dWg), ptr(dWb), ptr(db), ptr(ws), cur_stream()), "bwd_bits")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, fn in (("fwd", fwd), ("bwd", bwd), ("fwd, bit-packed mask", fwd_b), ("bwd, bit-packed mask", bwd_b)):
        for _ in range(5):
            fn()
        ts = []
        for _ in range(20):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        print("%s: median %.1f us  min %.1f us  (N = %d, L2 flushed between calls)" % (name, ts[len(ts) // 2], ts[0], N))


if __name__ == "__main__":
    main()
