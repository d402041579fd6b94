"""torch.autograd bindings + evaluation wrappers over the C ABI.

These make the CUDA kernels usable from the reference's own model code
(``total_loss.backward()`` in utility_train/trainer.py:55 just works); the fused
trainer path in ``idgrec.engine`` calls the same C entry points without autograd.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, cur_stream, ptr


def _f32c(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


class PropagateFn(torch.autograd.Function):
    """K x torch.sparse.mm + layer mean (LightGCN.py:36-52 / SimGCL.py:39-60 / XSimGCL.py:40-67)."""

    @staticmethod
    def forward(ctx, X0, graph, K, include_layer0, noise, eps, cl_layer):
        X0 = _f32c(X0)
        ctx.graph, ctx.K, ctx.inc0, ctx.cl = graph, K, include_layer0, cl_layer
        res = graph.propagate_fwd(X0, K, include_layer0, noise=noise, eps=eps, cl_layer=cl_layer)
        if cl_layer > 0:
            return res[0], res[1]
        return res

    @staticmethod
    def backward(ctx, gF, gC=None):
        gF = _f32c(gF)
        gC = _f32c(gC) if (ctx.cl > 0 and gC is not None) else None
        g = ctx.graph.propagate_bwd(gF, ctx.K, ctx.inc0, Gcl=gC, cl_layer=ctx.cl if gC is not None else 0)
        return g, None, None, None, None, None, None


def propagate(X0, graph, K, include_layer0=True, noise=None, eps=0.0, cl_layer=0):
    return PropagateFn.apply(X0, graph, K, include_layer0, noise, eps, cl_layer)


class SpmmFn(torch.autograd.Function):
    """One torch.sparse.mm(self.Graph, X) (NGCF.py:85); the adjacency is symmetric, so backward is the same product."""

    @staticmethod
    def forward(ctx, X, graph):
        X = _f32c(X)
        ctx.graph = graph
        Y = torch.empty_like(X)
        graph.spmm_layer(X, Y=Y)
        return Y

    @staticmethod
    def backward(ctx, gY):
        gY = _f32c(gY)
        gX = torch.empty_like(gY)
        ctx.graph.spmm_layer(gY, Y=gX)
        return gX, None


def spmm(X, graph):
    return SpmmFn.apply(X, graph)


class RowRangeSpmmFn(torch.autograd.Function):
    """Y[rows of g_out] = A_hat[those rows, :] . X, zero elsewhere -- e.g. torch.sparse.mm(R, item_embedding) of
    models/EGCF.py:52 with R the user rows of the symmetric bipartite matrix and X = [0; item_embedding].  Backward for a
    symmetric bipartite A_hat: the rows that can receive gradient are g_in's (the other side), gX = A_hat[g_in rows, :] . gY."""

    @staticmethod
    def forward(ctx, X, g_out, g_in):
        X = _f32c(X)
        ctx.g_in = g_in
        Y = torch.zeros_like(X)
        g_out.spmm_layer(X, Y=Y)
        return Y

    @staticmethod
    def backward(ctx, gY):
        gY = _f32c(gY)
        gX = torch.zeros_like(gY)
        ctx.g_in.spmm_layer(gY, Y=gX)
        return gX, None, None


def spmm_rows(X, g_out, g_in):
    return RowRangeSpmmFn.apply(X, g_out, g_in)


class TanhFn(torch.autograd.Function):
    """nn.Tanh (models/EGCF.py:42) on the CUDA library's element-wise kernels."""

    @staticmethod
    def forward(ctx, x):
        x = _f32c(x)
        y = torch.empty_like(x)
        check(_lib.lib().idg_tanh_fwd(ptr(x), ptr(y), x.numel(), cur_stream()), "idg_tanh_fwd")
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        gy = _f32c(gy)
        gx = torch.empty_like(gy)
        check(_lib.lib().idg_tanh_bwd(ptr(y), ptr(gy), ptr(gx), y.numel(), cur_stream()), "idg_tanh_bwd")
        return gx


def tanh(x):
    return TanhFn.apply(x)


class NgcfLayerFn(torch.autograd.Function):
    """One NGCF layer (NGCF.py:85-106): SpMM + fused dense epilogue kernels, forward and backward.
    Returns (D, O): the next layer's input and the row-normalised block of the final concat."""

    @staticmethod
    def forward(ctx, E, Wg, bg, Wb, bb, graph, keep, drop_p):
        l = _lib.lib()
        E, Wg, bg, Wb, bb = (_f32c(t) for t in (E, Wg, bg, Wb, bb))
        keep = _f32c(keep) if keep is not None else None
        N = E.shape[0]
        side = torch.empty_like(E)
        graph.spmm_layer(E, Y=side)
        S, D, O = torch.empty_like(E), torch.empty_like(E), torch.empty_like(E)
        check(l.idg_ngcf_dense_fwd(ptr(E), ptr(side), ptr(Wg), ptr(bg), ptr(Wb), ptr(bb), ptr(keep), float(drop_p), N, ptr(S), ptr(D), ptr(O), 64,
                                   cur_stream()), "idg_ngcf_dense_fwd")
        ctx.save_for_backward(E, side, S, D, Wg, Wb, keep if keep is not None else torch.empty(0, device=E.device))
        ctx.graph, ctx.drop_p, ctx.has_keep = graph, float(drop_p), keep is not None
        return D, O

    @staticmethod
    def backward(ctx, gD, gO):
        l = _lib.lib()
        E, side, S, D, Wg, Wb, keep = ctx.saved_tensors
        N = E.shape[0]
        gO = _f32c(gO) if gO is not None else torch.zeros_like(E)
        gD = _f32c(gD) if gD is not None else None
        dside, dEd = torch.empty_like(E), torch.empty_like(E)
        dWg, dWb = torch.empty_like(Wg), torch.empty_like(Wb)
        db = torch.empty(64, dtype=torch.float32, device=E.device)
        ws = torch.empty(int(l.idg_ngcf_workspace_bytes()), dtype=torch.uint8, device=E.device)
        check(l.idg_ngcf_dense_bwd(ptr(E), ptr(side), ptr(Wg), ptr(Wb), ptr(keep) if ctx.has_keep else None, ctx.drop_p, ptr(S), ptr(D), ptr(gO), 64,
                                   ptr(gD), N, ptr(dside), ptr(dEd), ptr(dWg), ptr(dWb), ptr(db), ptr(ws), cur_stream()), "idg_ngcf_dense_bwd")
        dE = torch.empty_like(E)
        ctx.graph.spmm_layer(dside, Y=dE, addend=dEd)   # dE = dE_direct + A_hat . dside (A_hat symmetric)
        return dE, dWg, db.view(1, 64), dWb, db.view(1, 64).clone(), None, None, None


def ngcf_layer(E, Wg, bg, Wb, bb, graph, keep=None, drop_p=0.0):
    return NgcfLayerFn.apply(E, Wg, bg, Wb, bb, graph, keep, drop_p)


class BprRegLossFn(torch.autograd.Function):
    """[bpr, reg_lambda*reg] = fused LightGCN.py:57-70 + losses.py:4-21."""

    @staticmethod
    def forward(ctx, F, E0, users, pos, neg, num_users, reg_lambda, reg_mask):
        l = _lib.lib()
        F, E0 = _f32c(F), _f32c(E0)
        users, pos, neg = (t.long().contiguous() for t in (users, pos, neg))
        B, (N, d) = int(users.numel()), F.shape
        ws = torch.empty(int(l.idg_bpr_workspace_bytes(B)), dtype=torch.uint8, device=F.device)
        loss = torch.empty(2, dtype=torch.float32, device=F.device)
        check(l.idg_bpr_forward(ptr(F), ptr(E0), ptr(users), ptr(pos), ptr(neg), B, num_users, N, d, float(reg_lambda),
                                int(reg_mask), ptr(loss), ptr(ws), cur_stream()), "idg_bpr_forward")
        ctx.save_for_backward(F, E0)
        ctx.ws, ctx.B, ctx.reg_lambda, ctx.reg_mask = ws, B, float(reg_lambda), int(reg_mask)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        l = _lib.lib()
        F, E0 = ctx.saved_tensors
        N, d = F.shape
        up = _f32c(gloss)
        G = torch.zeros_like(F)
        gE0 = torch.zeros_like(E0)
        check(l.idg_bpr_backward(ptr(F), ctx.B, d, ctx.reg_mask, ptr(up), ptr(G), 0.0, None, ptr(ctx.ws), cur_stream()), "idg_bpr_backward")
        check(l.idg_bpr_finish(ptr(E0), ptr(gE0), None, ctx.B, d, ctx.reg_lambda, ptr(up), None, ptr(ctx.ws), cur_stream()), "idg_bpr_finish")
        return G, gE0, None, None, None, None, None, None


def bpr_reg_loss(F, E0, users, pos, neg, num_users, reg_lambda, reg_mask=7):
    return BprRegLossFn.apply(F, E0, users, pos, neg, num_users, reg_lambda, reg_mask)


class InfoNCEFn(torch.autograd.Function):
    """losses.get_InfoNCE_loss(V1[idx], V2[idx], tau) (losses.py:24-35) on rows idx of two [N,64] views."""

    @staticmethod
    def forward(ctx, V1, V2, idx, temperature):
        l = _lib.lib()
        V1, V2 = _f32c(V1), _f32c(V2)
        idx = idx.long().contiguous()
        n, d = int(idx.numel()), V1.shape[1]
        ws = torch.empty(int(l.idg_infonce_workspace_bytes(n, d)), dtype=torch.uint8, device=V1.device)
        loss = torch.zeros(1, dtype=torch.float32, device=V1.device)
        g1, g2 = torch.zeros_like(V1), torch.zeros_like(V2)
        check(l.idg_infonce_fwd_bwd(ptr(V1), ptr(V2), ptr(idx), n, d, float(temperature), 1.0, ptr(loss), ptr(g1), ptr(g2),
                                    ptr(ws), cur_stream()), "idg_infonce_fwd_bwd")
        ctx.save_for_backward(g1, g2)
        return loss[0]

    @staticmethod
    def backward(ctx, gl):
        g1, g2 = ctx.saved_tensors
        return g1 * gl, g2 * gl, None, None


def infonce_rows(V1, V2, idx, temperature):
    return InfoNCEFn.apply(V1, V2, idx, temperature)


class GatherRowsFn(torch.autograd.Function):
    """``table[idx]`` (e.g. all_user_embeddings[user.long()], models/LightCCF.py:66) as a dense [n,d] block; the
    backward scatters with duplicates summed in entry order by their first occurrence (no float atomics)."""

    @staticmethod
    def forward(ctx, T, idx):
        T = _f32c(T)
        idx = idx.long().contiguous()
        n, d = int(idx.numel()), T.shape[1]
        out = torch.empty((n, d), dtype=torch.float32, device=T.device)
        check(_lib.lib().idg_gather_rows(ptr(T), ptr(idx), n, d, ptr(out), cur_stream()), "idg_gather_rows")
        ctx.save_for_backward(idx)
        ctx.shape = T.shape
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = _f32c(g)
        gT = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        check(_lib.lib().idg_scatter_add_rows(ptr(g), ptr(idx), int(idx.numel()), g.shape[1], ptr(gT), cur_stream()), "idg_scatter_add_rows")
        return gT, None


def gather_rows(T, idx):
    return GatherRowsFn.apply(T, idx)


PAIR_KINDS = {"lightccf": 0, "lightcscf": 1, "sccf_down": 2, "sccf_up": 3, "align": 4, "uniform": 5}


class PairLossFn(torch.autograd.Function):
    """One batch x batch loss of the LightGCN-backbone models (idg_pair_loss; include/idgrec.h) on dense [n,d]
    blocks: forward and both gradients come out of one call, the backward only scales them."""

    @staticmethod
    def forward(ctx, X, Y, kind, p0, p1):
        l = _lib.lib()
        X = _f32c(X)
        Y = _f32c(Y) if Y is not None else None
        n, d = X.shape
        ws = torch.empty(int(l.idg_pair_loss_workspace_bytes(n, d)), dtype=torch.uint8, device=X.device)
        loss = torch.empty(1, dtype=torch.float32, device=X.device)
        need = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        gX = torch.empty_like(X) if need else None
        gY = torch.empty_like(Y) if (need and Y is not None) else None
        check(l.idg_pair_loss(int(kind), ptr(X), ptr(Y), n, d, float(p0), float(p1), ptr(loss), ptr(gX), ptr(gY), ptr(ws), cur_stream()),
              "idg_pair_loss")
        ctx.has_y = Y is not None
        ctx.save_for_backward(*[t for t in (gX, gY) if t is not None])
        return loss[0]

    @staticmethod
    def backward(ctx, gl):
        saved = ctx.saved_tensors
        gX = saved[0] * gl
        gY = saved[1] * gl if (ctx.has_y and len(saved) > 1) else None
        return gX, gY, None, None, None


def pair_loss(kind, X, Y=None, p0=0.0, p1=0.0):
    """kind: one of PAIR_KINDS.  X, Y: [n,d] (un-normalised) rows; returns a scalar tensor."""
    return PairLossFn.apply(X, Y, PAIR_KINDS[kind], p0, p1)


# ---------------------------------------------------------------------------------------
# evaluation
# ---------------------------------------------------------------------------------------

def eval_topk(Fu, Fi, users, mask_indptr, mask_indices, K, want_scores=False, ws=None):
    """ids [nu,K] int64 (and exact scores) of the top-K unmasked items per user, (score desc, id asc)."""
    l = _lib.lib()
    Fu, Fi = _f32c(Fu), _f32c(Fi)
    users = users.long().contiguous()
    nu, (U, d), I = int(users.numel()), Fu.shape, Fi.shape[0]
    if ws is None:
        ws = torch.empty(int(l.idg_eval_workspace_bytes(nu, I, d, K)), dtype=torch.uint8, device=Fu.device)
    ids = torch.empty((nu, K), dtype=torch.int64, device=Fu.device)
    sc = torch.empty((nu, K), dtype=torch.float32, device=Fu.device) if want_scores else None
    check(l.idg_eval_topk(ptr(Fu), ptr(Fi), U, I, d, ptr(mask_indptr), ptr(mask_indices), ptr(users), nu, K, ptr(ids), ptr(sc),
                          ptr(ws), cur_stream()), "idg_eval_topk")
    return (ids, sc) if want_scores else ids


def rating_matrix(Fu, Fi, users):
    """sigmoid(Fu[users] . Fi^T) as a dense [b, I] tensor (LightGCN.py:74-80); API completeness only."""
    Fu, Fi = _f32c(Fu), _f32c(Fi)
    users = users.long().contiguous()
    out = torch.empty((users.numel(), Fi.shape[0]), dtype=torch.float32, device=Fu.device)
    check(_lib.lib().idg_rating_matrix(ptr(Fu), ptr(Fi), ptr(users), users.numel(), Fi.shape[0], Fu.shape[1], ptr(out), cur_stream()), "idg_rating_matrix")
    return out


def eval_metric_sums(topk_ids, users, test_indptr, test_indices, ks, ws=None):
    """float64 [len(ks), 3] sums over users of (recall, precision, ndcg) at each k (metrics.py:4-36)."""
    l = _lib.lib()
    nu, K = topk_ids.shape
    nk = len(ks)
    if ws is None:
        ws = torch.empty(int(l.idg_eval_workspace_bytes(nu, 1, 64, K)), dtype=torch.uint8, device=topk_ids.device)
    sums = torch.empty(3 * nk, dtype=torch.float64, device=topk_ids.device)
    ks_arr = (C.c_int32 * nk)(*[int(k) for k in ks])
    check(l.idg_eval_metrics(ptr(topk_ids.contiguous()), ptr(users.long().contiguous()), nu, K, ptr(test_indptr), ptr(test_indices),
                             ks_arr, nk, ptr(sums), ptr(ws), cur_stream()), "idg_eval_metrics")
    return sums.view(nk, 3)


def adam_step(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    check(_lib.lib().idg_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
                                   int(step), cur_stream()), "idg_adam_step")


def neg_sample_replay(train_user, pos_indptr, pos_indices, cand):
    """Exact replay of data_loader.py:108-127 against a bulk numpy candidate stream (host)."""
    l = _lib.lib()
    tu = np.ascontiguousarray(train_user, dtype=np.int64)
    ip = np.ascontiguousarray(pos_indptr, dtype=np.int32)
    ix = np.ascontiguousarray(pos_indices, dtype=np.int32)
    cand = np.ascontiguousarray(cand, dtype=np.int64)
    neg = np.empty(len(tu), dtype=np.int64)
    used = C.c_int64(0)
    rc = l.idg_neg_sample_replay(tu.ctypes.data, len(tu), ip.ctypes.data, ix.ctypes.data, cand.ctypes.data, len(cand),
                                 neg.ctypes.data, C.byref(used))
    if rc == -2:
        return None, int(used.value)
    check(rc, "idg_neg_sample_replay")
    return neg, int(used.value)


def parse_ratings(path):
    """data_loader.py:48-70 on the host in one pass (idg_parse_ratings): (line_user, line_len, users, items,
    max_user, max_item); users/items in file order, one entry per interaction."""
    l = _lib.lib()
    n_pairs, n_lines, mu, mi = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
    bp = str(path).encode()
    check(l.idg_parse_ratings(bp, None, None, 0, C.byref(n_pairs), None, None, 0, C.byref(n_lines), C.byref(mu), C.byref(mi)), "idg_parse_ratings")
    users, items = np.empty(n_pairs.value, dtype=np.int64), np.empty(n_pairs.value, dtype=np.int64)
    line_user, line_len = np.empty(n_lines.value, dtype=np.int64), np.empty(n_lines.value, dtype=np.int64)
    if n_pairs.value or n_lines.value:
        check(l.idg_parse_ratings(bp, users.ctypes.data, items.ctypes.data, max(len(users), 1), C.byref(n_pairs), line_user.ctypes.data,
                                  line_len.ctypes.data, max(len(line_user), 1), C.byref(n_lines), C.byref(mu), C.byref(mi)), "idg_parse_ratings")
    return line_user, line_len, users, items, int(mu.value), int(mi.value)
