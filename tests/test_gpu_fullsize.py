"""Size-independent properties at the full amazon-book shape (BASELINE.json configs[1]: 52,643 users / 91,599 items /
2.38 M train edges, synthetic), where the CPU oracle is too slow to run whole: CSR invariants, SpMM linearity /
self-adjointness / row sums / determinism, restricted == full propagation on the batch rows, top-K ordering and
exactness on a sample of users (fp64 on the device), metrics bounds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ab():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from idgrec import datagen
    from idgrec.graph import Graph, build_norm_adjacency
    dev = torch.device("cuda:0")
    g = datagen.gen_graph("amazon-book")
    csr = build_norm_adjacency(g.train_user, g.train_item, g.num_users, g.num_items, device=dev)
    return dev, g, csr, Graph(csr)


def test_csr_invariants_full_size(ab):
    dev, g, csr, G = ab
    U, I = g.num_users, g.num_items
    N = U + I
    ip, ix, dt = csr.indptr.long(), csr.indices.long(), csr.data
    assert int(ip[0]) == 0 and int(ip[-1]) == ix.numel() == 2 * len(g.train_user)
    deg = ip[1:] - ip[:-1]
    assert bool((deg >= 0).all())
    rows = torch.repeat_interleave(torch.arange(N, device=dev), deg)
    # bipartite: user rows point at item columns and vice versa; columns strictly ascending inside a row
    assert bool(((rows < U) == (ix >= U)).all())
    same_row = rows[1:] == rows[:-1]
    assert bool((ix[1:][same_row] > ix[:-1][same_row]).all())
    # symmetric structure and bitwise symmetric values: sort the transposed entries and compare
    key, keyT = rows * N + ix, ix * N + rows
    order = torch.argsort(keyT)
    assert torch.equal(keyT[order], key)
    assert torch.equal(dt[order].view(torch.int32), dt.view(torch.int32))
    # value rule (data_graph.py:46-51 in float32): data = (d[r] * 1) * d[c] with d = deg^-0.5
    d = csr.dinv
    assert torch.equal(((d[rows] * 1.0) * d[ix]).view(torch.int32), dt.view(torch.int32))


def test_spmm_properties_full_size(ab):
    dev, g, csr, G = ab
    N = csr.shape[0]
    gen = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn(N, 64, generator=gen, device=dev)
    Y = torch.randn(N, 64, generator=gen, device=dev)
    AX, AY, AZ = torch.empty_like(X), torch.empty_like(X), torch.empty_like(X)
    G.spmm_layer(X, Y=AX)
    G.spmm_layer(Y, Y=AY)
    # determinism: bit-identical on repetition
    AX2 = torch.empty_like(X)
    G.spmm_layer(X, Y=AX2)
    assert torch.equal(AX, AX2)
    # linearity
    Z = (0.5 * X - 2.0 * Y).contiguous()
    G.spmm_layer(Z, Y=AZ)
    ref = 0.5 * AX - 2.0 * AY
    assert float((AZ - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    # self-adjoint (A_hat symmetric): <A X, Y> == <X, A Y>
    a, b = float((AX.double() * Y.double()).sum()), float((X.double() * AY.double()).sum())
    scale = float((AX.double().abs() * Y.double().abs()).sum())           # the terms cancel: compare against their total magnitude
    assert abs(a - b) <= 1e-6 * scale
    # row sums: A_hat . 1 equals the segmented sum of the CSR values (fp64 reference on the device)
    ones = torch.ones(N, 64, device=dev)
    A1 = torch.empty_like(ones)
    G.spmm_layer(ones, Y=A1)
    rows = torch.repeat_interleave(torch.arange(N, device=dev), (csr.indptr[1:] - csr.indptr[:-1]).long())
    rs = torch.zeros(N, dtype=torch.float64, device=dev).index_add_(0, rows, csr.data.double())
    assert float((A1[:, 0].double() - rs).abs().max()) <= 1e-5 * float(rs.abs().max())
    assert torch.equal(A1[:, 0], A1[:, 63])


def test_restricted_equals_full_and_bwd_chain_full_size(ab):
    from idgrec.graph import BatchRows
    dev, g, csr, G = ab
    U, I = g.num_users, g.num_items
    N, B, K = U + I, 1024, 3
    gen = torch.Generator(device=dev).manual_seed(1)
    X0 = (torch.rand(N, 64, generator=gen, device=dev) - 0.5) * 0.1
    e = torch.randint(0, len(g.train_user), (B,), generator=gen, device=dev).cpu().numpy()
    u = torch.from_numpy(g.train_user[e]).to(dev)
    p = torch.from_numpy(g.train_item[e]).to(dev)
    n = torch.randint(0, I, (B,), generator=gen, device=dev)
    rows = BatchRows(N, B, dev)
    rows.build(u.data_ptr(), p.data_ptr(), n.data_ptr(), B, U)
    touched = torch.unique(torch.cat([u, U + p, U + n]))
    assert int(rows.count.item()) == touched.numel()
    full = G.propagate_fwd(X0, K, True)
    part = G.propagate_fwd(X0, K, True, rows=rows)
    assert torch.equal(full[touched], part[touched])
    Gd = torch.zeros(N, 64, device=dev)
    Gd[touched] = torch.randn(touched.numel(), 64, generator=gen, device=dev) * 1e-3
    dense = G.propagate_bwd(Gd, K, True)
    sparse = G.propagate_bwd(Gd, K, True, rows=rows)
    assert float((dense - sparse).abs().max()) <= 1e-6 * float(dense.abs().max())
    # backward is the adjoint of forward: <P X0, Gd> == <X0, P^T Gd>
    a, b = float((full.double() * Gd.double()).sum()), float((X0.double() * dense.double()).sum())
    scale = float((full.double().abs() * Gd.double().abs()).sum())
    assert abs(a - b) <= 1e-5 * scale
    rows.clear()


@pytest.mark.parametrize("scale", [0.05, 0.4])
def test_topk_properties_full_size(ab, scale):
    """Ordering, masking and exactness of the tensor-core ranking at 52,643 x 91,599 (exact check on 512 sampled users)."""
    from idgrec import ops
    import scipy.sparse as sp
    dev, g, csr, G = ab
    U, I, K = g.num_users, g.num_items, 20
    net = sp.csr_matrix((np.ones(len(g.train_user)), (g.train_user, g.train_item)), shape=(U, I))
    net.sort_indices()
    mp = torch.from_numpy(net.indptr.astype(np.int32)).to(dev)
    mi = torch.from_numpy(net.indices.astype(np.int32)).to(dev)
    gen = torch.Generator(device=dev).manual_seed(2)
    Fu = torch.randn(U, 64, generator=gen, device=dev) * scale
    Fi = torch.randn(I, 64, generator=gen, device=dev) * scale
    users = torch.arange(U, device=dev)
    ids, sc = ops.eval_topk(Fu, Fi, users, mp, mi, K, want_scores=True)
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())                          # scores descending
    srt = torch.sort(ids, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                        # ids distinct
    assert bool(((ids >= 0) & (ids < I)).all())
    # none of the returned ids is a train positive: look each (user, id) pair up in the sorted mask keys
    key = torch.from_numpy((g.train_user * I + g.train_item).astype(np.int64)).to(dev)
    key = torch.sort(key).values
    q = (users[:, None] * I + ids).reshape(-1)
    pos = torch.searchsorted(key, q).clamp_(max=key.numel() - 1)
    assert not bool((key[pos] == q).any())
    # exact rank on a sample: fp64 scores on the device, masked, stable (score desc, id asc)
    samp = torch.randperm(U, generator=gen, device=dev)[:512]
    S = Fu[samp].double() @ Fi.double().t()
    for j, uu in enumerate(samp.tolist()):
        S[j, net.indices[net.indptr[uu]:net.indptr[uu + 1]]] = -float("inf")
    ref = torch.sort(-S, dim=1, stable=True).indices[:, :K]
    assert torch.equal(ids[samp], ref)


# ---------------------------------------------------------------- the scale-up shape (BASELINE.json configs[4])
def test_xl_shape_properties():
    """1M users x 1M items x 100M edges (nnz 200M), generated on the device like bench.py's xl record.  The CPU oracle cannot
    run at this size, so parity rests on size-independent properties: CSR structure, sampled rows of one propagation layer
    against an fp64 evaluation straight from the CSR, both closure constructions agreeing bit for bit, and the row-partitioned
    trainer (one partition, neighbourhood restriction on) reproducing the dense single-GPU fused trainer bit for bit."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~40 GB of device memory")
    from idgrec import _lib, datagen
    from idgrec._lib import check, cur_stream, ptr
    from idgrec.dist import DistFusedTrainer
    from idgrec.engine import FusedTrainer
    from idgrec.graph import BatchRows, Graph, build_norm_adjacency
    dev = torch.device("cuda:0")
    U = I = 1000000
    E = 100000000
    N = U + I
    eu, ei = datagen.gen_edges_device(U, I, E, 2024, dev)
    csr = build_norm_adjacency(eu, ei, U, I, device=dev)
    ip, ix = csr.indptr.long(), csr.indices
    assert csr.nnz == 2 * E and int(ip[0]) == 0 and int(ip[-1]) == 2 * E
    assert bool((ip[1:] >= ip[:-1]).all())
    assert int(ix[: int(ip[U])].min()) >= U and int(ix[int(ip[U]):].max()) < U          # bipartite blocks
    G = Graph(csr)
    gen = torch.Generator(device=dev).manual_seed(3)
    X = (torch.rand(N, 64, generator=gen, device=dev) - 0.5) * 0.1
    Y = torch.empty_like(X)
    G.spmm_layer(X, Y=Y)
    rows = torch.randint(0, N, (256,), generator=gen, device=dev)
    for r in rows.tolist()[:256]:
        s, e = int(ip[r]), int(ip[r + 1])
        want = (csr.data[s:e].double()[:, None] * X[ix[s:e].long()].double()).sum(0)
        got = Y[r].double()
        assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-12, r
    # closure of a batch: pass over all nonzeros (idg_closure_bitmap) == union of the batch rows' neighbour lists (idg_closure_from_rows)
    B = 1024
    sel = torch.randint(0, E, (B,), generator=gen, device=dev)
    bu, bp, bn = eu[sel].contiguous(), ei[sel].contiguous(), torch.randint(0, I, (B,), generator=gen, device=dev)
    br = BatchRows(N, B, dev)
    br.build(ptr(bu), ptr(bp), ptr(bn), B, U)
    c1, c2 = torch.zeros_like(br.bitmap), torch.zeros_like(br.bitmap)
    l = _lib.lib()
    check(l.idg_closure_bitmap(G._h, ptr(br.bitmap), ptr(c1), cur_stream()), "idg_closure_bitmap")
    check(l.idg_closure_from_rows(G._h, ptr(br.rowlist), ptr(br.count), br.max_rows, ptr(br.bitmap), ptr(c2), cur_stream()), "idg_closure_from_rows")
    assert torch.equal(c1, c2)
    frac = float(sum(bin(w & 0xffffffff).count("1") for w in c1.cpu().tolist())) / N
    assert 0.05 < frac < 0.9, frac
    del Y, c1, c2, br
    # row-partitioned trainer with one partition and the neighbourhood restriction vs the dense single-GPU trainer
    table = (torch.rand(N, 64, generator=gen, device=dev) * 2 - 1) * (6.0 / (U + 64)) ** 0.5
    ref = FusedTrainer("LightGCN", G, table.clone(), U, 3, 1e-4, 1e-3, max_batch=B, use_cuda_graph=False, restrict_rows=False)
    ft = DistFusedTrainer("LightGCN", csr, table.clone(), U, 3, 1e-4, 1e-3, 0, 1, max_batch=B, use_cuda_graph=True, full_graph=G, closure_restrict=True)
    del table
    for step in range(3):
        sel = torch.randint(0, E, (B,), generator=gen, device=dev)
        b = (eu[sel].contiguous(), ei[sel].contiguous(), torch.randint(0, I, (B,), generator=gen, device=dev))
        l0, l1 = ref.step(*b).clone(), ft.step(*b).clone()
        assert torch.equal(l0, l1), (step, l0, l1)
    assert torch.equal(ref.E0, ft.E0)
