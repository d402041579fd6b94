"""The real incumbent: the reference's LightGCN path run ON THE GPU through torch / cuSPARSE / cuBLAS.

TEST INFRASTRUCTURE / REPORTED BASELINE ONLY -- like everything under oracle/, this file is never imported by the
product (id-grec_b200/); only bench.py's baseline legs call it.  It restates, statement by statement, what the
reference executes when ``main.py`` finds a CUDA device (BASELINE.md section 3 "optional but recommended"):

  * graph handle: ``convert_sp_mat_to_sp_tensor(...).coalesce().to(device)`` -- a coalesced COO tensor
    (utility/utility_function/tools.py:95-109, models/LightGCN.py:30-32);
  * train step: ``aggregate`` = cat -> K x ``torch.sparse.mm`` -> stack -> mean -> split (LightGCN.py:36-52), BPR + L2
    (LightGCN.py:54-72, losses.py:4-21), ``Optim.zero_grad(); loss.backward(); Optim.step()`` with
    ``torch.optim.Adam`` and one ``.item()`` per loss (utility_train/trainer.py:40-56);
  * evaluation batch: ``get_rating_for_test`` re-runs ``aggregate`` and materialises sigmoid(U_b . F_i^T)
    (LightGCN.py:74-80); Python lists of the batch's train positives, ``rating[rows, cols] = -1``, ``torch.topk``,
    ``.cpu()`` (utility_train/batch_test.py:52-70).

Nothing here is a kernel of this repository: it is the library path the hand-written kernels have to beat.
"""
from __future__ import annotations

import time

import numpy as np
import torch


def _sync(dev):
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)


class ReferenceLightGCNOnDevice(torch.nn.Module):
    """models/LightGCN.py:14-80 on ``device`` (same parameters, same ops, same order)."""

    def __init__(self, A_coo: torch.Tensor, user_w: np.ndarray, item_w: np.ndarray, K: int, reg_lambda: float, device):
        super().__init__()
        self.U, self.I, self.K, self.reg_lambda = user_w.shape[0], item_w.shape[0], K, reg_lambda
        self.user_embedding = torch.nn.Embedding.from_pretrained(torch.from_numpy(user_w), freeze=False)
        self.item_embedding = torch.nn.Embedding.from_pretrained(torch.from_numpy(item_w), freeze=False)
        self.Graph = A_coo.coalesce().to(device)
        self.activation = torch.nn.Sigmoid()
        self.to(device)

    def aggregate(self):
        all_embedding = torch.cat([self.user_embedding.weight, self.item_embedding.weight])
        embeddings = [all_embedding]
        for _ in range(self.K):
            all_embedding = torch.sparse.mm(self.Graph, all_embedding)
            embeddings.append(all_embedding)
        final = torch.mean(torch.stack(embeddings, dim=1), dim=1)
        return torch.split(final, [self.U, self.I])

    def forward(self, user, positive, negative):
        users, items = self.aggregate()
        u, p, n = users[user.long()], items[positive.long()], items[negative.long()]
        ego_u, ego_p, ego_n = self.user_embedding(user), self.item_embedding(positive), self.item_embedding(negative)
        pos = torch.sum(u * p, dim=1)
        neg = torch.sum(u * n, dim=1)
        bpr = torch.mean(-torch.log(torch.sigmoid(pos - neg) + 10e-8))
        reg = 0
        for e in (ego_u, ego_p, ego_n):
            reg = reg + 0.5 * e.norm(2).pow(2) / float(e.shape[0])
        return [bpr, self.reg_lambda * reg]

    def get_rating_for_test(self, user):
        users, items = self.aggregate()
        return self.activation(torch.matmul(users[user.long()], items.t()))


def epoch_estimate(g, device, n_train_batches=40, n_eval_batches=4, batch=1024, test_batch=1024, K=3, reg_lambda=1e-4, lr=1e-3, top_k=20):
    """Seconds per epoch of the reference's own GPU path on synthetic graph ``g`` (idgrec.datagen.SynthGraph): a bounded
    sample of train steps and evaluation batches, wall clock with the device synchronised, extrapolated to the epoch's
    batch counts (sampling / shuffling on the host excluded, exactly like the CPU baseline)."""
    import scipy.sparse as sp
    from oracle import ref_oracle as O
    dev = torch.device(device)
    U, I, E = g.num_users, g.num_items, len(g.train_user)
    net = sp.csr_matrix((np.ones(E), (g.train_user, g.train_item)), shape=(U, I))
    net.sort_indices()
    ip, ix, dt, _ = O.norm_adjacency(net)
    A = O.csr_to_torch_coo(ip, ix, dt, U + I)
    gen = torch.Generator().manual_seed(2024)
    model = ReferenceLightGCNOnDevice(A, O.xavier_uniform(U, 64, gen), O.xavier_uniform(I, 64, gen), K, reg_lambda, dev)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    rng = np.random.default_rng(1)

    def one_step():
        e = rng.integers(0, E, batch)
        bu = torch.from_numpy(g.train_user[e]).to(dev)
        bp = torch.from_numpy(g.train_item[e]).to(dev)
        bn = torch.from_numpy(rng.integers(0, I, batch)).to(dev)
        loss_list = model(bu, bp, bn)
        total = sum(loss_list)
        opt.zero_grad()
        total.backward()
        opt.step()
        return [float(l.item()) for l in loss_list]      # trainer.py:52: one sync per loss

    model.train()
    for _ in range(5):
        one_step()
    _sync(dev)
    t0 = time.perf_counter()
    for _ in range(n_train_batches):
        one_step()
    _sync(dev)
    t_batch = (time.perf_counter() - t0) / n_train_batches

    users = np.unique(g.test_user)
    indptr, indices = net.indptr, net.indices
    model.eval()

    def one_eval(b):
        bu = users[b * test_batch:(b + 1) * test_batch]
        with torch.no_grad():
            rating = model.get_rating_for_test(torch.from_numpy(bu).to(dev))
            exclude_index, exclude_items = [], []                      # batch_test.py:54-63: Python lists
            for r, u in enumerate(bu):
                items = indices[indptr[u]:indptr[u + 1]]
                exclude_index.extend([r] * len(items))
                exclude_items.extend(items)
            rating[exclude_index, exclude_items] = -1
            _, top = torch.topk(rating, k=top_k)
            return top.cpu()

    one_eval(0)
    _sync(dev)
    t0 = time.perf_counter()
    for b in range(n_eval_batches):
        one_eval(b)
    _sync(dev)
    t_eval = (time.perf_counter() - t0) / n_eval_batches
    nb = (E + batch - 1) // batch
    neb = (len(users) + test_batch - 1) // test_batch
    return {"value": t_batch * nb + t_eval * neb, "unit": "s/epoch", "kind": "reference ops on the GPU (torch.sparse.mm / autograd / torch.optim.Adam / matmul + topk)",
            "t_train_batch_ms": t_batch * 1e3, "t_eval_batch_ms": t_eval * 1e3, "train_batches": nb, "eval_batches": neb,
            "sample": "%d train batches of %d and %d eval batches of %d users, wall clock with device sync, extrapolated to %d + %d batches; host sampling excluded"
                      % (n_train_batches, batch, n_eval_batches, test_batch, nb, neb),
            "torch": torch.__version__}
