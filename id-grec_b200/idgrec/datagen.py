"""Synthetic interaction graphs of the reference's dataset shapes.

The reference mount ships only ``test.txt`` for yelp2018 / amazon-book (its
``train.txt`` blobs are missing, ``.MISSING_LARGE_BLOBS``), so every measurement
and parity test runs on synthetic graphs with the published shapes
(SURVEY.md section 8 d): user degree ~ lognormal(3.3, 0.9) with a floor, item
popularity ~ lognormal(0, 1.1), unique (user, item) pairs, every user has at
least one train and one test item, ~79/21 split.  Files are written in the
reference's format (``user item item ...`` per line,
utility/utility_data/data_loader.py:48-70) so they load through ``Data``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

SHAPES = {
    # name: (users, items, train interactions, test interactions)
    "tiny": (120, 150, 2600, 700),
    "small": (2000, 3000, 60000, 16000),
    "medium": (8000, 12000, 300000, 80000),
    "yelp2018": (31668, 38048, 1237259, 324147),
    "amazon-book": (52643, 91599, 2380730, 603382),
    "xl": (1000000, 1000000, 100000000, 1000000),
}


@dataclass
class SynthGraph:
    num_users: int
    num_items: int
    train_user: np.ndarray  # int64, grouped by user (file order)
    train_item: np.ndarray
    test_user: np.ndarray
    test_item: np.ndarray


def _pairs(U, I, total, rng, sigma_u=0.9, sigma_i=1.1, floor=2):
    wu = rng.lognormal(3.3, sigma_u, U)
    wi = rng.lognormal(0.0, sigma_i, I)
    pu = wu / wu.sum()
    pi = wi / wi.sum()
    cu, ci = np.cumsum(pu), np.cumsum(pi)
    cu[-1] = ci[-1] = 1.0
    keys = np.empty(0, dtype=np.int64)
    # guaranteed `floor` distinct items per user
    base_u = np.repeat(np.arange(U, dtype=np.int64), floor)
    base_i = (rng.integers(0, I, U)[:, None] + np.arange(floor)[None, :] * max(1, I // (floor + 1))) % I
    keys = base_u * I + base_i.reshape(-1)
    need = total
    while True:
        n = int((need - len(np.unique(keys))) * 1.08) + 1024
        u = np.searchsorted(cu, rng.random(n)).astype(np.int64)
        i = np.searchsorted(ci, rng.random(n)).astype(np.int64)
        keys = np.unique(np.concatenate([keys, u * I + i]))
        if len(keys) >= need:
            break
    if len(keys) > need:
        # drop random surplus pairs but never the guaranteed ones
        base = np.unique(base_u * I + base_i.reshape(-1))
        extra = np.setdiff1d(keys, base, assume_unique=True)
        keep = rng.permutation(len(extra))[: need - len(base)]
        keys = np.sort(np.concatenate([base, extra[keep]]))
    return keys // I, keys % I


def gen_graph(shape="tiny", seed=2024, test_frac=None) -> SynthGraph:
    if isinstance(shape, str):
        U, I, ntr, nte = SHAPES[shape]
    else:
        U, I, ntr, nte = shape
    rng = np.random.default_rng(seed)
    u, i = _pairs(U, I, ntr + nte, rng)
    # per-user random order, then the first part of every user's list is train
    r = rng.random(len(u))
    order = np.lexsort((r, u))
    u, i = u[order], i[order]
    deg = np.bincount(u, minlength=U)
    start = np.concatenate([[0], np.cumsum(deg)[:-1]])
    rank = np.arange(len(u)) - start[u]
    frac = ntr / float(ntr + nte) if test_frac is None else 1.0 - test_frac
    ntrain_u = np.clip(np.rint(frac * deg).astype(np.int64), 1, np.maximum(deg - 1, 1))
    # fix the total to hit ntr exactly where possible
    diff = int(ntr - ntrain_u.sum())
    if diff != 0:
        room = (deg - 1 - ntrain_u) if diff > 0 else (ntrain_u - 1)
        cand = np.flatnonzero(room > 0)
        pick = rng.permutation(cand)[: abs(diff)]
        ntrain_u[pick] += 1 if diff > 0 else -1
    is_train = rank < ntrain_u[u]
    return SynthGraph(U, I, u[is_train], i[is_train], u[~is_train], i[~is_train])


def _write(path, users, items, U):
    deg = np.bincount(users, minlength=U)
    ptr = np.concatenate([[0], np.cumsum(deg)])
    s = items.astype(str)
    with open(path, "w") as f:
        out = []
        for uu in range(U):
            a, b = ptr[uu], ptr[uu + 1]
            if b > a:
                out.append(str(uu) + " " + " ".join(s[a:b]) + "\n")
        f.write("".join(out))


def write_dataset(root: str, name: str, g: SynthGraph) -> str:
    """Writes ``root/name/{train,test}.txt`` and returns the directory."""
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    _write(os.path.join(d, "train.txt"), g.train_user, g.train_item, g.num_users)
    _write(os.path.join(d, "test.txt"), g.test_user, g.test_item, g.num_users)
    return d


def gen_edges_device(num_users: int, num_items: int, num_edges: int, seed: int = 2024, device="cuda"):
    """Scale-up graphs (SURVEY.md section 8 d, XL: 1M x 1M, 100M edges) generated directly on the device as
    sorted unique (user, item) pairs -- same degree families as gen_graph (lognormal user activity sigma 0.9,
    item popularity sigma 1.1), never through text or scipy.  Deterministic for a given seed and GPU model, so
    every rank of a multi-GPU run builds the identical graph."""
    import torch
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    wu = torch.exp(torch.randn(num_users, generator=gen, device=dev, dtype=torch.float64) * 0.9 + 3.3)
    wi = torch.exp(torch.randn(num_items, generator=gen, device=dev, dtype=torch.float64) * 1.1)
    cu = torch.cumsum(wu / wu.sum(), 0)
    ci = torch.cumsum(wi / wi.sum(), 0)
    keys = (torch.arange(num_users, device=dev, dtype=torch.int64) * num_items
            + torch.randint(0, num_items, (num_users,), generator=gen, device=dev))      # every user has an edge
    while True:
        need = num_edges - keys.numel()
        if need <= 0:
            break
        n = int(need * 1.1) + 4096
        u = torch.searchsorted(cu, torch.rand(n, generator=gen, device=dev, dtype=torch.float64)).clamp_(max=num_users - 1)
        i = torch.searchsorted(ci, torch.rand(n, generator=gen, device=dev, dtype=torch.float64)).clamp_(max=num_items - 1)
        keys = torch.unique(torch.cat([keys, u * num_items + i]))
        del u, i
    if keys.numel() > num_edges:
        keep = torch.randperm(keys.numel(), generator=gen, device=dev)[:num_edges]
        keys = torch.sort(keys[keep]).values
    return keys // num_items, keys % num_items
