"""CPU oracle for the ID-GRec hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
path (``id-grec_b200/``) never does: it fails loudly when the CUDA library is
missing instead of falling back to anything in here.

What it is: a restatement of the reference's algorithm for the path
graph-build -> propagate -> BPR/InfoNCE -> full-ranking eval, written against the
same third-party arithmetic the reference uses (torch CPU ``sparse.mm`` /
``matmul`` / autograd / ``optim.Adam``, scipy, numpy legacy RNG).  Every function
cites the reference file:line it follows.  It is *pinned* by
``tests/golden/*.npz`` -- outputs of the unmodified reference modules imported
from ``/root/reference`` in the build container by
``tests/golden/make_golden.py`` (the reference ships no tests or golden vectors
of its own) -- see ``tests/test_oracle_golden.py``.

Third-party pins of the reference (README.md:10-16, no lockfile): python 3.8.18,
pytorch 2.1.0, scipy 1.10.1, numpy 1.24.3.  Here: torch 2.11, scipy 1.18,
numpy 2.3; the properties relied on were re-verified.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
import torch

# --------------------------------------------------------------------------
# a1  data loading  (utility/utility_data/data_loader.py:27-70,129-133,151-159)
# --------------------------------------------------------------------------


@dataclass
class OracleData:
    path: str
    num_users: int
    num_items: int
    num_nodes: int
    num_train: int
    num_test: int
    train_user: np.ndarray
    train_item: np.ndarray
    test_user: np.ndarray
    test_item: np.ndarray
    user_item_net: sp.csr_matrix
    all_positive: List[np.ndarray]
    test_dict: Dict[int, List[int]]
    config: Optional[dict] = None
    split_test_dict: Optional[list] = None

    def get_user_pos_items(self, users):
        # data_loader.py:129-133
        return [self.all_positive[u] for u in users]


def read_ratings(file_name: str):
    """data_loader.py:48-70.  One line = ``user item item ...``; users whose line
    has no item are recorded in ``unique`` but contribute nothing else (and do
    not move ``max_user``)."""
    users, items, unique = [], [], []
    max_user, max_item = 0, 0
    with open(file_name, "r") as f:
        for line in f:
            if line == "":
                break
            tok = [int(t) for t in line.strip().split(" ")]
            u, pos = tok[0], tok[1:]
            unique.append(u)
            if not pos:
                continue
            max_user = max(max_user, u)
            max_item = max(max_item, max(pos))
            users.extend([u] * len(pos))
            items.extend(pos)
    return (np.array(unique), np.array(users, dtype=np.int64), np.array(items, dtype=np.int64),
            len(items), max_user, max_item)


def load_dataset(path: str, config: Optional[dict] = None) -> OracleData:
    """data_loader.py:27-46 (load_data) + :151-159 (build_test)."""
    _, tu, ti, ntrain, mu1, mi1 = read_ratings(path + "/train.txt")
    _, su, si, ntest, mu2, mi2 = read_ratings(path + "/test.txt")
    U = max(mu1, mu2) + 1
    I = max(mi1, mi2) + 1
    # duplicates SUM (weight 2.0) exactly like csr_matrix((ones,(u,i))) -- data_loader.py:42-43
    net = sp.csr_matrix((np.ones(len(tu)), (tu, ti)), shape=(U, I))
    net.sum_duplicates()
    net.sort_indices()
    all_pos = [net.indices[net.indptr[u]:net.indptr[u + 1]] for u in range(U)]
    test_dict: Dict[int, List[int]] = {}
    for u, i in zip(su.tolist(), si.tolist()):
        test_dict.setdefault(u, []).append(i)
    return OracleData(path, U, I, U + I, ntrain, ntest, tu, ti, su, si, net, all_pos, test_dict, config)


# --------------------------------------------------------------------------
# a2  normalised adjacency  (utility/utility_data/data_graph.py:7-55)
# --------------------------------------------------------------------------


def norm_adjacency(user_item_net: sp.csr_matrix, add_self: bool = False):
    """Canonical CSR (indptr int32, indices int32 ascending per row, data fp32) of
    ``D^-1/2 [[0,R],[R^T,0]] D^-1/2`` (+I before normalising when ``add_self``).

    Arithmetic rule (verified against the reference by the
    golden vectors):
      * no-self (data_graph.py:33-55): everything in float32:
        ``deg = rowsum_f32``, ``d = np.power(deg, -0.5)`` in float32, inf->0,
        ``data = (d[row] * a) * d[col]`` in float32 (a = 1.0, or 2.0 for a
        duplicated pair).
      * with-self (data_graph.py:7-30): ``dok_f32 + sp.eye`` promotes to float64:
        ``d = np.power(deg_f64, -0.5)``, ``data64 = (d[row]*a)*d[col]``, rounded to
        fp32 only by tools.py:101.
    """
    R = user_item_net.tocoo()
    U, I = R.shape
    N = U + I
    rows = np.concatenate([R.row, R.col + U]).astype(np.int64)
    cols = np.concatenate([R.col + U, R.row]).astype(np.int64)
    if add_self:
        w = np.concatenate([R.data, R.data]).astype(np.float64)
        rows = np.concatenate([rows, np.arange(N)])
        cols = np.concatenate([cols, np.arange(N)])
        w = np.concatenate([w, np.ones(N)])
        deg = np.bincount(rows, weights=w, minlength=N).astype(np.float64)
        with np.errstate(divide="ignore"):
            d = np.power(deg, -0.5)
        d[np.isinf(d)] = 0.0
        data = ((d[rows] * w) * d[cols]).astype(np.float32)
    else:
        w = np.concatenate([R.data, R.data]).astype(np.float32)
        deg = np.bincount(rows, weights=w.astype(np.float64), minlength=N).astype(np.float32)
        with np.errstate(divide="ignore"):
            d = np.power(deg, np.float32(-0.5)).astype(np.float32)
        d[np.isinf(d)] = 0.0
        data = ((d[rows] * w).astype(np.float32) * d[cols]).astype(np.float32)
    order = np.lexsort((cols, rows))
    rows, cols, data = rows[order], cols[order], data[order]
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    return indptr, cols.astype(np.int32), data, d


def csr_to_torch_coo(indptr, indices, data, N) -> torch.Tensor:
    """tools.py:95-109 + ``.coalesce()`` (LightGCN.py:31-32): coalesced fp32 COO."""
    rows = np.repeat(np.arange(N, dtype=np.int64), np.diff(indptr.astype(np.int64)))
    idx = torch.from_numpy(np.stack([rows, indices.astype(np.int64)]))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(data.astype(np.float32)), (N, N)).coalesce()


# --------------------------------------------------------------------------
# a4 / a5  sampler, shuffle, batching  (data_loader.py:108-127; tools.py:35-64)
# --------------------------------------------------------------------------


def sample_negatives(data: OracleData) -> np.ndarray:
    """data_loader.py:108-127, one np.random.randint draw at a time exactly as the
    reference does (slow; use on small graphs).  Consumes the numpy global RNG."""
    out = []
    for e in range(len(data.train_user)):
        u = data.train_user[e]
        pos = data.all_positive[u]
        if len(pos) == 0:
            continue
        while True:
            neg = np.random.randint(0, data.num_items)
            if neg in pos:
                continue
            break
        out.append([u, data.train_item[e], neg])
    return np.array(out)


def sample_negatives_bulk(data: OracleData) -> np.ndarray:
    """Same stream, same result, same final RNG state as ``sample_negatives`` using
    the bulk-replay identity of SURVEY.md section 8 a4: scalar ``randint(0,n)`` calls
    == one ``randint(0,n,size=k)``; walk the candidate stream skipping candidates
    that are positives of the *current* edge's user."""
    E = len(data.train_user)
    st = np.random.get_state()
    slack = max(1024, E // 50)
    while True:
        np.random.set_state(st)
        cand = np.random.randint(0, data.num_items, size=E + slack)
        neg = np.empty(E, dtype=np.int64)
        j = 0
        ok = True
        indptr, indices = data.user_item_net.indptr, data.user_item_net.indices
        for e in range(E):
            u = data.train_user[e]
            lo, hi = indptr[u], indptr[u + 1]
            while True:
                if j >= len(cand):
                    ok = False
                    break
                c = cand[j]
                j += 1
                k = np.searchsorted(indices[lo:hi], c)
                if k < hi - lo and indices[lo + k] == c:
                    continue
                neg[e] = c
                break
            if not ok:
                break
        if ok:
            break
        slack *= 4
    np.random.set_state(st)
    np.random.randint(0, data.num_items, size=j)  # leave the stream where the reference would
    return np.stack([data.train_user, data.train_item, neg], axis=1)


def shuffle_indices(n: int) -> np.ndarray:
    """tools.py:41-42."""
    idx = np.arange(n)
    np.random.shuffle(idx)
    return idx


def mini_batches(n: int, batch_size: int):
    """tools.py:55-64: slices, last one short."""
    for s in range(0, n, batch_size):
        yield s, min(n, s + batch_size)


# --------------------------------------------------------------------------
# a6 / a7  propagation  (models/LightGCN.py:36-52, SimGCL.py:39-60, XSimGCL.py:40-67)
# --------------------------------------------------------------------------


def propagate(A: torch.Tensor, X0: torch.Tensor, K: int, include_layer0: bool,
              noises: Optional[Sequence[torch.Tensor]] = None, eps: float = 0.0,
              cl_layer: int = 0):
    """K x ``torch.sparse.mm`` then mean over the stacked layer outputs.

    ``include_layer0`` True  -> LightGCN (LightGCN.py:41,47-48)
    ``include_layer0`` False -> SimGCL / XSimGCL (SimGCL.py:45,55-56)
    ``noises`` (K tensors of U[0,1) shaped like X) switches on the perturbation
    ``x += sign(x) * normalize(noise, dim=-1) * eps`` (SimGCL.py:49-51); the noise
    is injected, not drawn, so both sides of a parity test see the same values.
    ``cl_layer`` > 0 also returns the post-noise output of layer index cl_layer-1
    (XSimGCL.py:57-58)."""
    x = X0
    outs = [x] if include_layer0 else []
    x_cl = X0
    for layer in range(K):
        x = torch.sparse.mm(A, x)
        if noises is not None:
            x = x + torch.sign(x) * torch.nn.functional.normalize(noises[layer], dim=-1) * eps
        outs.append(x)
        if layer == cl_layer - 1:
            x_cl = x
    final = torch.mean(torch.stack(outs, dim=1), dim=1)
    return (final, x_cl) if cl_layer > 0 else final


# --------------------------------------------------------------------------
# a9 / a10  losses  (utility/utility_function/losses.py:4-35)
# --------------------------------------------------------------------------


def bpr_loss(u, p, n):
    """losses.py:4-13 (note 10e-8 == 1e-7)."""
    pos = torch.sum(u * p, dim=1)
    neg = torch.sum(u * n, dim=1)
    return torch.mean(-torch.log(torch.sigmoid(pos - neg) + 10e-8))


def reg_loss(*embs):
    """losses.py:16-21: sum_t 1/2 ||E_t||^2 / rows(E_t)."""
    r = 0
    for e in embs:
        r = r + 0.5 * e.norm(2).pow(2) / float(e.shape[0])
    return r


def infonce_loss(e1, e2, temperature):
    """losses.py:24-35 (note 10e-6 == 1e-5)."""
    e1 = torch.nn.functional.normalize(e1)
    e2 = torch.nn.functional.normalize(e2)
    pos = torch.exp((e1 * e2).sum(dim=-1) / temperature)
    ttl = torch.exp(torch.matmul(e1, e2.t()) / temperature).sum(dim=1)
    return torch.mean(-torch.log(pos / ttl + 10e-6))


# --------------------------------------------------------------------------
# whole training step (models/*.py forward + trainer.py:42-56)
# --------------------------------------------------------------------------


@dataclass
class StepResult:
    losses: List[float]
    grad_user: np.ndarray
    grad_item: np.ndarray
    extra: dict = field(default_factory=dict)


class OracleModel:
    """fp32 torch-CPU model with the reference's forward semantics for
    LightGCN / SimGCL / XSimGCL (MFBPR = K 0 without graph)."""

    def __init__(self, kind: str, A: Optional[torch.Tensor], user_w: np.ndarray, item_w: np.ndarray,
                 K: int = 3, reg_lambda: float = 1e-4, ssl_lambda: float = 0.0, eps: float = 0.0,
                 temperature: float = 0.2, cl_layer: int = 1, lr: float = 1e-3):
        assert kind in ("LightGCN", "SimGCL", "XSimGCL", "MFBPR")
        self.kind, self.A, self.K = kind, A, K
        self.reg_lambda, self.ssl_lambda, self.eps = reg_lambda, ssl_lambda, eps
        self.temperature, self.cl_layer = temperature, cl_layer
        self.user_w = torch.nn.Parameter(torch.from_numpy(np.array(user_w, dtype=np.float32)))
        self.item_w = torch.nn.Parameter(torch.from_numpy(np.array(item_w, dtype=np.float32)))
        self.U, self.I = self.user_w.shape[0], self.item_w.shape[0]
        self.opt = torch.optim.Adam([self.user_w, self.item_w], lr=lr)  # trainer.py:11

    def aggregate(self, noises=None, want_cl=False):
        X0 = torch.cat([self.user_w, self.item_w])
        if self.kind == "MFBPR":
            return self.user_w, self.item_w
        if self.kind == "LightGCN":
            F = propagate(self.A, X0, self.K, True)
            return torch.split(F, [self.U, self.I])
        if want_cl:
            F, C = propagate(self.A, X0, self.K, False, noises, self.eps, self.cl_layer)
            return (*torch.split(F, [self.U, self.I]), *torch.split(C, [self.U, self.I]))
        F = propagate(self.A, X0, self.K, False, noises, self.eps)
        return torch.split(F, [self.U, self.I])

    def forward(self, user, pos, neg, noises=None):
        """LightGCN.py:54-72 / SimGCL.py:62-90 / XSimGCL.py:69-95.
        ``noises``: SimGCL -> [view1 K tensors, view2 K tensors]; XSimGCL -> K tensors."""
        user, pos, neg = (torch.as_tensor(t, dtype=torch.long) for t in (user, pos, neg))
        ssl = None
        if self.kind == "XSimGCL":
            fu, fi, cu, ci = self.aggregate(noises, want_cl=True)
        else:
            fu, fi = self.aggregate()
        bpr = bpr_loss(fu[user], fi[pos], fi[neg])
        reg = self.reg_lambda * reg_loss(self.user_w[user], self.item_w[pos], self.item_w[neg])
        if self.kind in ("SimGCL", "XSimGCL"):
            uidx, iidx = torch.unique(user), torch.unique(pos)
            if self.kind == "SimGCL":
                u1, i1 = self.aggregate(noises[0])
                u2, i2 = self.aggregate(noises[1])
                ssl_u = infonce_loss(u1[uidx], u2[uidx], self.temperature)
                ssl_i = infonce_loss(i1[iidx], i2[iidx], self.temperature)
            else:
                ssl_u = infonce_loss(cu[uidx], fu[uidx], self.temperature)
                ssl_i = infonce_loss(ci[iidx], fi[iidx], self.temperature)
            ssl = self.ssl_lambda * (ssl_u + ssl_i)
        return [bpr, reg] + ([ssl] if ssl is not None else [])

    def step(self, user, pos, neg, noises=None, apply_adam=True) -> StepResult:
        """trainer.py:42-56: sum the losses, zero_grad, backward, Adam step."""
        losses = self.forward(user, pos, neg, noises)
        total = 0.0
        for l in losses:
            total = total + l
        self.opt.zero_grad()
        total.backward()
        gu, gi = self.user_w.grad.detach().numpy().copy(), self.item_w.grad.detach().numpy().copy()
        if apply_adam:
            self.opt.step()
        return StepResult([float(l.item()) for l in losses], gu, gi)

    @torch.no_grad()
    def final_embeddings(self):
        fu, fi = self.aggregate()[:2]
        return fu.detach().numpy().copy(), fi.detach().numpy().copy()


# --------------------------------------------------------------------------
# a8  NGCF  (models/NGCF.py:67-111)  -- dropout masks are injected
# --------------------------------------------------------------------------


def ngcf_aggregate(A, X0, W_gcn, b_gcn, W_bi, b_bi, keep_masks=None, drop_p=None):
    """NGCF.py:83-108.  ``keep_masks[l]`` is a {0,1} tensor like E (the reference
    draws it inside nn.Dropout, always in training mode -- SURVEY.md section 3.4);
    kept entries are scaled by 1/(1-p)."""
    ego = X0
    outs = [ego]
    for l in range(len(W_gcn)):
        side = torch.sparse.mm(A, ego)
        s = torch.matmul(side, W_gcn[l]) + b_gcn[l]
        bi = torch.matmul(ego * side, W_bi[l]) + b_bi[l]
        ego = torch.nn.functional.leaky_relu(s + bi, negative_slope=0.2)
        if keep_masks is not None:
            ego = ego * keep_masks[l] / (1.0 - drop_p[l])
        outs.append(torch.nn.functional.normalize(ego, p=2, dim=1))
    return torch.cat(outs, dim=1)


# --------------------------------------------------------------------------
# a13 / a14  full-ranking evaluation
# --------------------------------------------------------------------------


def scores_fp64_sequential(Fu: np.ndarray, Fi: np.ndarray) -> np.ndarray:
    """Exact-rank score definition (SURVEY.md section 7 tier T0): fp64 dot product of the
    fp32 embeddings accumulated in index order k = 0..d-1.  Products of two fp32
    values are exact in fp64, so FMA vs mul+add cannot differ; only the order
    matters and it is fixed.  The CUDA rescore uses the same order."""
    a = Fu.astype(np.float64)
    b = Fi.astype(np.float64)
    acc = np.zeros((a.shape[0], b.shape[0]), dtype=np.float64)
    for k in range(a.shape[1]):
        acc += a[:, k:k + 1] * b[None, :, k]
    return acc


def topk_exact(Fu, Fi, users, mask_indptr, mask_indices, K: int) -> Tuple[np.ndarray, np.ndarray]:
    """T0 oracle: train positives removed (LightGCN.py:74-80 + batch_test.py:62-68),
    order = score descending then item id ascending.  Returns ids [n,K] int64 and
    fp64 scores [n,K]."""
    users = np.asarray(users, dtype=np.int64)
    S = scores_fp64_sequential(Fu[users], Fi)
    for r, u in enumerate(users):
        S[r, mask_indices[mask_indptr[u]:mask_indptr[u + 1]]] = -np.inf
    ids = np.argsort(-S, axis=1, kind="stable")[:, :K]  # stable => ties keep ascending id
    return ids.astype(np.int64), np.take_along_axis(S, ids, axis=1)


def topk_reference_faithful(Fu, Fi, users, mask_indptr, mask_indices, K: int):
    """T1 oracle: the reference's own arithmetic -- fp32 matmul -> fp32 sigmoid ->
    ``rating[eu, ei] = -1`` (batch_test.py:65) -- but with a *defined* tie order (id
    ascending) where torch.topk's is arbitrary.  Also returns the fp32 rating."""
    users = np.asarray(users, dtype=np.int64)
    rating = torch.sigmoid(torch.matmul(torch.from_numpy(Fu[users]), torch.from_numpy(Fi).t())).numpy()
    for r, u in enumerate(users):
        rating[r, mask_indices[mask_indptr[u]:mask_indptr[u + 1]]] = -1.0
    ids = np.argsort(-rating, axis=1, kind="stable")[:, :K]
    return ids.astype(np.int64), rating


# --------------------------------------------------------------------------
# a15  metrics  (utility/utility_function/metrics.py:4-58, batch_test.py:80-107)
# --------------------------------------------------------------------------


def hit_matrix(topk_ids: np.ndarray, truth: Sequence[Sequence[int]]) -> np.ndarray:
    """metrics.py:49-58."""
    r = np.zeros(topk_ids.shape, dtype=np.float64)
    for i, t in enumerate(truth):
        r[i] = np.isin(topk_ids[i], np.asarray(list(t)))
    return r


def metric_sums(r: np.ndarray, truth_len: np.ndarray, k: int):
    """Sums over the users of recall@k, precision@k, ndcg@k (metrics.py:4-36)."""
    hits = r[:, :k].sum(1)
    recall = float(np.sum(hits / truth_len))
    precision = float(np.sum(hits) / k)
    disc = 1.0 / np.log2(np.arange(2, k + 2))
    ideal = (np.arange(k)[None, :] < np.minimum(truth_len, k)[:, None]).astype(np.float64)
    idcg = (ideal * disc).sum(1)
    dcg = (r[:, :k] * disc).sum(1)
    idcg[idcg == 0.0] = 1.0
    nd = dcg / idcg
    nd[np.isnan(nd)] = 0.0
    return recall, precision, float(nd.sum())


def evaluate(Fu, Fi, data: OracleData, top_K: Sequence[int], test_batch_size: int, mode: str = "exact"):
    """batch_test.py:37-93 with the propagation hoisted out (its result does not
    depend on the batch).  Sums are accumulated per test batch in the reference's
    order (float64) and divided by the number of test users."""
    users = list(data.test_dict.keys())
    indptr, indices = data.user_item_net.indptr, data.user_item_net.indices
    res = {k: np.zeros(len(top_K)) for k in ("precision", "recall", "hit", "ndcg")}
    fn = topk_exact if mode == "exact" else topk_reference_faithful
    all_ids = []
    for s, e in mini_batches(len(users), test_batch_size):
        bu = users[s:e]
        ids, _ = fn(Fu, Fi, bu, indptr, indices, max(top_K))
        all_ids.append(ids)
        truth = [data.test_dict[u] for u in bu]
        r = hit_matrix(ids, truth)
        tl = np.array([len(t) for t in truth], dtype=np.float64)
        for j, k in enumerate(top_K):
            rc, pr, nd = metric_sums(r, tl, k)
            res["recall"][j] += rc
            res["precision"][j] += pr
            res["ndcg"][j] += nd
    for k in ("recall", "precision", "ndcg"):
        res[k] /= float(len(users))
    return res, np.concatenate(all_ids)


# --------------------------------------------------------------------------
# section 8 f rows: activity split, sparsity test, SGL sub-graphs, loss-only models
# --------------------------------------------------------------------------


def sparsity_split(data: OracleData):
    """data_loader.py:161-204, statement by statement (``count`` is never advanced there, so every
    group closes at 25 % of all interactions; the remainder -- possibly empty -- closes at the end)."""
    by_level: Dict[int, List[int]] = {}
    for uid in data.test_dict.keys():
        by_level.setdefault(len(data.all_positive[uid]) + len(data.test_dict[uid]), []).append(uid)
    total = data.num_train + data.num_test
    groups, temp, n_rates, n_count = [], [], 0, total
    levels = sorted(by_level)
    for idx, level in enumerate(levels):
        temp = temp + by_level[level]
        n_rates += level * len(by_level[level])
        n_count -= level * len(by_level[level])
        if n_rates >= 0.25 * total:
            groups.append(temp)
            temp, n_rates = [], 0
        if idx == len(levels) - 1 or n_count == 0:
            groups.append(temp)
    return groups


def evaluate_groups(Fu, Fi, data: OracleData, groups, top_K: Sequence[int], mode: str = "exact"):
    """batch_test.py:110-170: Test()'s metrics per activity group, each divided by the group size."""
    indptr, indices = data.user_item_net.indptr, data.user_item_net.indices
    fn = topk_exact if mode == "exact" else topk_reference_faithful
    out = []
    for users in groups:
        ids, _ = fn(Fu, Fi, list(users), indptr, indices, max(top_K))
        truth = [data.test_dict[u] for u in users]
        r = hit_matrix(ids, truth)
        tl = np.array([len(t) for t in truth], dtype=np.float64)
        res = {k: np.zeros(len(top_K)) for k in ("precision", "recall", "hit", "ndcg")}
        for j, k in enumerate(top_K):
            rc, pr, nd = metric_sums(r, tl, k)
            res["recall"][j], res["precision"][j], res["ndcg"][j] = rc / len(users), pr / len(users), nd / len(users)
        out.append(res)
    return out


def subgraph_adjacency(user_item_net: sp.csr_matrix, keep_index: np.ndarray):
    """tools.py:67-92 (``ed``/``rw``) for a given list of kept edges (the reference draws it with the
    unseeded python ``random.sample``): float32 ones, A = G + G^T, d = rowsum^-1/2 (inf -> 0) in
    float32, D A D.  Canonical CSR (indptr, indices, data fp32)."""
    U, I = user_item_net.shape
    N = U + I
    ui, ii = user_item_net.nonzero()
    ui, ii = np.array(ui)[keep_index], np.array(ii)[keep_index]
    g = sp.csr_matrix((np.ones_like(ui, dtype=np.float32), (ui, ii + U)), shape=(N, N))
    adj = g + g.T
    with np.errstate(divide="ignore"):
        d_inv = np.power(np.array(adj.sum(axis=1)), -0.5).flatten()
    d_inv[np.isinf(d_inv)] = 0.0
    dm = sp.diags(d_inv)
    out = dm.dot(adj).dot(dm).tocsr()
    out.sort_indices()
    return out.indptr, out.indices, out.data


def lightccf_na_loss(e1, e2, tau):
    """models/LightCCF.py:81-94 (10e-6 == 1e-5)."""
    e1, e2 = torch.nn.functional.normalize(e1), torch.nn.functional.normalize(e2)
    pos = torch.exp((e1 * e2).sum(dim=-1) / tau)
    ttl = torch.exp((torch.matmul(e1, e2.t()) + torch.matmul(e1, e1.t())) / tau).sum(dim=1)
    return torch.mean(-torch.log(pos / ttl + 10e-6))


def lightcscf_loss(e1, e2, tau, margin):
    """models/LightCSCF.py:93-104."""
    e1, e2 = torch.nn.functional.normalize(e1, dim=1), torch.nn.functional.normalize(e2, dim=1)
    sim = (e1 * e2).sum(dim=-1)
    pos = torch.exp(sim / tau) + torch.exp(torch.relu(sim - margin) / tau)
    t = torch.matmul(e1, e2.t()) + torch.matmul(e1, e1.t())
    ttl = (torch.exp(t / tau) + torch.exp(torch.relu(t - margin) / tau)).sum(dim=1)
    return torch.mean(-torch.log(pos / ttl + 10e-6))


def sccf_losses(fu, fi, user, pos, tau):
    """models/SCCF.py:59-81 -> [-up, down]."""
    u_idx, u_counts = torch.unique(user, return_counts=True)
    i_idx, i_counts = torch.unique(pos, return_counts=True)
    u_counts, i_counts = u_counts.reshape(-1, 1).float(), i_counts.reshape(-1, 1).float()
    a, b = torch.nn.functional.normalize(fu[user], dim=-1), torch.nn.functional.normalize(fi[pos], dim=-1)
    ip = (a * b).sum(dim=1)
    up = ((ip / tau).exp() + (ip ** 2 / tau).exp()).log().mean()
    a, b = torch.nn.functional.normalize(fu[u_idx], dim=-1), torch.nn.functional.normalize(fi[i_idx], dim=-1)
    sim = a @ b.t()
    score = (sim / tau).exp() + (sim ** 2 / tau).exp()
    down = (score * (u_counts @ i_counts.t())).mean().log()
    return [-up, down]


def align_loss(e1, e2):
    """losses.py:61-64."""
    e1, e2 = torch.nn.functional.normalize(e1, dim=-1), torch.nn.functional.normalize(e2, dim=-1)
    return torch.mean((e1 - e2).norm(p=2, dim=1).pow(2))


def uniform_loss(e):
    """losses.py:67-69."""
    e = torch.nn.functional.normalize(e, dim=-1)
    return torch.pdist(e, p=2).pow(2).mul(-2).exp().mean().log()


def next_model_step(kind: str, A, user_w, item_w, user, pos, neg, cfg: dict, encoder: str = "LightGCN", sub_graphs=None, K: int = 3):
    """forward + backward of LightCCF / LightCSCF / SCCF / DirectAU / SGL on one batch
    (models/LightCCF.py:58-79, LightCSCF.py:58-91, SCCF.py:54-81, DirectAU.py:59-79, SGL.py:60-89)."""
    uw = torch.nn.Parameter(torch.from_numpy(np.array(user_w, dtype=np.float32)))
    iw = torch.nn.Parameter(torch.from_numpy(np.array(item_w, dtype=np.float32)))
    U, I = uw.shape[0], iw.shape[0]
    user, pos, neg = (torch.as_tensor(t, dtype=torch.long) for t in (user, pos, neg))

    def agg(graph):
        return torch.split(propagate(graph, torch.cat([uw, iw]), K, True), [U, I])

    fu, fi = (uw, iw) if encoder == "MF" else agg(A)
    ue, pe, ne = fu[user], fi[pos], fi[neg]
    reg3 = reg_loss(uw[user], iw[pos], iw[neg])
    if kind == "LightCCF":
        losses = [bpr_loss(ue, pe, ne), float(cfg["reg_lambda"]) * reg3,
                  float(cfg["ssl_lambda"]) * lightccf_na_loss(ue, pe, float(cfg["temperature"]))]
    elif kind == "LightCSCF":
        na = float(cfg["lambda_gamma"]) * lightcscf_loss(ue, pe, float(cfg["temperature"]), float(cfg["lambda_margin"]))
        reg = float(cfg["lambda_reg"]) * reg3
        losses = [bpr_loss(ue, pe, ne), reg, na] if encoder == "MF" else [reg, na]
    elif kind == "SCCF":
        losses = sccf_losses(fu, fi, user, pos, float(cfg["temperature"]))
    elif kind == "DirectAU":
        losses = [align_loss(ue, pe), float(cfg["gamma"]) * (uniform_loss(ue) + uniform_loss(pe)) / 2,
                  float(cfg["reg_lambda"]) * reg_loss(uw[user], iw[pos])]
    elif kind == "SGL":
        u1, i1 = agg(sub_graphs[0])
        u2, i2 = agg(sub_graphs[1])
        tau = float(cfg["temperature"])
        ssl = infonce_loss(u1[user], u2[user], tau) + infonce_loss(i1[pos], i2[pos], tau)
        losses = [bpr_loss(ue, pe, ne), float(cfg["reg_lambda"]) * reg3, float(cfg["ssl_lambda"]) * ssl]
    else:
        raise ValueError(kind)
    total = 0.0
    for l in losses:
        total = total + l
    total.backward()
    return StepResult([float(l.item()) for l in losses], uw.grad.numpy().copy(), iw.grad.numpy().copy())


def bipartite_adjacency(user_item_net: sp.csr_matrix):
    """data_graph.py:56-77: D_u^-1/2 R D_i^-1/2 in float64 (user_item_net holds float64 ones), as the fp32 torch COO
    tensor the model holds after tools.py:95-109 + ``.coalesce()`` (models/EGCF.py:33-35)."""
    R = user_item_net
    with np.errstate(divide="ignore"):
        rd = np.power(np.array(R.sum(axis=1)), -0.5).flatten()
        cd = np.power(np.array(R.sum(axis=0)), -0.5).flatten()
    rd[np.isinf(rd)] = 0.0
    cd[np.isinf(cd)] = 0.0
    M = sp.diags(rd).dot(R).dot(sp.diags(cd)).tocoo().astype(np.float32)
    idx = torch.from_numpy(np.stack([M.row, M.col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(M.data), M.shape).coalesce()


def egcf_aggregate(Rt: torch.Tensor, A: Optional[torch.Tensor], item_w: torch.Tensor, K: int, mode: str):
    """models/EGCF.py:45-84: users = tanh(R items); 'parallel': K x tanh(A_hat .) on [users; items], summed;
    'alternating': users <- tanh(R items), items <- tanh(R^T users), each side summed over the layers."""
    U, I = Rt.shape
    if mode == "parallel":
        users = torch.tanh(torch.sparse.mm(Rt, item_w))
        x = torch.cat([users, item_w])
        outs = []
        for _ in range(K):
            x = torch.tanh(torch.sparse.mm(A, x))
            outs.append(x)
        return torch.split(torch.sum(torch.stack(outs, dim=1), dim=1), [U, I])
    items, us, its = item_w, [], []
    for _ in range(K):
        users = torch.tanh(torch.sparse.mm(Rt, items))
        items = torch.tanh(torch.sparse.mm(Rt.transpose(0, 1), users))
        us.append(users)
        its.append(items)
    return torch.sum(torch.stack(us, dim=1), dim=1), torch.sum(torch.stack(its, dim=1), dim=1)


def egcf_step(Rt, A, item_w, user, pos, neg, cfg: dict, mode: str, K: int = 3):
    """models/EGCF.py:86-112 forward + backward -> (losses, grad_item, final_users, final_items)."""
    iw = torch.nn.Parameter(torch.from_numpy(np.array(item_w, dtype=np.float32)))
    user, pos, neg = (torch.as_tensor(t, dtype=torch.long) for t in (user, pos, neg))
    fu, fi = egcf_aggregate(Rt, A, iw, K, mode)
    ue, pe, ne = fu[user], fi[pos], fi[neg]
    tau = float(cfg["temperature"])
    losses = [bpr_loss(ue, pe, ne), float(cfg["reg_lambda"]) * reg_loss(iw[pos], iw[neg]),
              float(cfg["ssl_lambda"]) * (infonce_loss(ue, ue, tau) + infonce_loss(pe, pe, tau) + infonce_loss(ue, pe, tau))]
    total = 0.0
    for l in losses:
        total = total + l
    total.backward()
    return [float(l.item()) for l in losses], iw.grad.numpy().copy(), fu.detach().numpy().copy(), fi.detach().numpy().copy()


# --------------------------------------------------------------------------
# timing helper for bench.py's cpu_baseline / --impl reference legs
# --------------------------------------------------------------------------


def xavier_uniform(rows: int, cols: int, gen: torch.Generator) -> np.ndarray:
    bound = math.sqrt(6.0 / (rows + cols))
    return ((torch.rand(rows, cols, generator=gen) * 2 - 1) * bound).numpy()
