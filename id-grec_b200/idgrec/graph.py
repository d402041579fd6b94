"""Device graph objects: normalised adjacency (CSR) and the propagation handle.

Host-side mirror of ``self.Graph`` (models/LightGCN.py:30-32): where the reference
holds a coalesced torch COO tensor and calls ``torch.sparse.mm`` on it, this holds
an ``idg_graph`` handle of libidgrec_sm100.so.  All arithmetic runs in the CUDA
library; numpy is used only for ``np.power(deg, -0.5)`` -- the same numpy call the
reference makes (data_graph.py:46), which keeps the edge weights bit-identical.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, cur_stream, ptr


class DeviceCSR:
    """Canonical CSR of the normalised adjacency on the device (int32 / int32 / fp32)."""

    def __init__(self, indptr, indices, data, dinv, n_rows, n_cols):
        self.indptr, self.indices, self.data, self.dinv = indptr, indices, data, dinv
        self.shape = (n_rows, n_cols)

    @property
    def nnz(self):
        return int(self.indices.numel())

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.data.cpu().numpy(), self.indices.cpu().numpy(), self.indptr.cpu().numpy()), shape=self.shape)


def build_norm_adjacency(train_user, train_item, num_users: int, num_items: int, add_self: bool = False,
                         device="cuda", f64_degrees: bool | None = None) -> DeviceCSR:
    """data_graph.py:33-55 (``add_self=False``) / :7-30 (``add_self=True``) on the device.
    ``f64_degrees`` (default: same as add_self): d = deg^-1/2 and the products in float64, rounded to fp32 once --
    the arithmetic of sparse_adjacency_matrix_with_self and of sparse_adjacency_matrix_R (data_graph.py:56-77, whose
    D_u^-1/2 R D_i^-1/2 is the upper-right block of this symmetric matrix and its transpose the lower-left one)."""
    f64 = add_self if f64_degrees is None else bool(f64_degrees)
    l = _lib.lib()
    dev = torch.device(device)
    if torch.is_tensor(train_user):
        u, i = train_user.to(dev, torch.int64).contiguous(), train_item.to(dev, torch.int64).contiguous()
    else:
        u = torch.as_tensor(np.ascontiguousarray(train_user), dtype=torch.int64).to(dev)
        i = torch.as_tensor(np.ascontiguousarray(train_item), dtype=torch.int64).to(dev)
    E = int(u.numel())
    N = num_users + num_items
    cap = 2 * E + (N if add_self else 0)
    with torch.cuda.device(dev):
        indptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
        indices = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        mult = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
        deg = torch.empty(N, dtype=torch.float64, device=dev)
        nnz = C.c_int64(0)
        check(l.idg_csr_structure(ptr(u), ptr(i), E, num_users, num_items, int(add_self), ptr(indptr), ptr(indices),
                                  ptr(mult), ptr(deg), C.byref(nnz), cur_stream()), "idg_csr_structure")
        nnz = int(nnz.value)
        deg_h = deg.cpu().numpy()
        with np.errstate(divide="ignore"):
            if f64:  # dok_f32 + sp.eye promotes to float64 (data_graph.py:19-24); user_item_net is float64 (:62-70)
                d = np.power(deg_h, -0.5)
            else:         # float32 throughout (data_graph.py:44-48)
                d = np.power(deg_h.astype(np.float32), np.float32(-0.5)).astype(np.float32)
        d[np.isinf(d)] = 0.0
        d_dev = torch.from_numpy(d).to(dev)
        data = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        check(l.idg_csr_normalise(ptr(indptr), ptr(indices), ptr(mult), N, nnz,
                                  None if f64 else ptr(d_dev), ptr(d_dev) if f64 else None, ptr(data), cur_stream()),
              "idg_csr_normalise")
        indices = indices[:nnz].clone()
        data = data[:nnz].clone()
    return DeviceCSR(indptr, indices, data, d_dev, N, N)


class Graph:
    """Propagation handle over rows [row_begin, row_end) of a DeviceCSR (whole matrix by default)."""

    def __init__(self, csr: DeviceCSR, row_begin: int = 0, row_end: int | None = None):
        l = _lib.lib()
        N = csr.shape[0]
        row_end = N if row_end is None else row_end
        self.N, self.row_begin, self.row_end = N, row_begin, row_end
        self.device = csr.indptr.device
        self.csr = csr
        with torch.cuda.device(self.device):
            if row_begin == 0 and row_end == N:
                indptr, indices, data = csr.indptr, csr.indices, csr.data
            else:
                ip = csr.indptr[row_begin:row_end + 1]
                s, e = int(ip[0].item()), int(ip[-1].item())
                indptr = (ip - ip[0]).contiguous()
                indices, data = csr.indices[s:e].contiguous(), csr.data[s:e].contiguous()
            self.nnz = int(indices.numel())
            h = C.c_void_p()
            check(l.idg_graph_create(ptr(indptr), ptr(indices) if self.nnz else None, ptr(data) if self.nnz else None,
                                     row_end - row_begin, csr.shape[1], self.nnz, row_begin, C.byref(h), cur_stream()),
                  "idg_graph_create")
        self._h = h
        self._work = {}

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().idg_graph_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def is_whole(self):
        return self.row_begin == 0 and self.row_end == self.N

    def work(self, d: int, tag: str = "w"):
        key = (d, tag)
        if key not in self._work:
            self._work[key] = torch.empty(2 * self.N * d, dtype=torch.float32, device=self.device)
        return self._work[key]

    def spmm_layer(self, X, Y=None, addend=None, noise=None, eps=0.0, acc_in=None, acc_out=None, acc_div=1.0):
        d = X.shape[1]
        check(_lib.lib().idg_spmm_layer(self._h, ptr(X), ptr(Y), ptr(addend), ptr(noise), float(eps), ptr(acc_in), ptr(acc_out),
                                        float(acc_div), d, cur_stream()), "idg_spmm_layer")

    def propagate_fwd(self, X0, K, include_layer0, noise=None, eps=0.0, cl_layer=0, out_mean=None, out_cl=None, rows=None):
        """K layers + layer mean.  ``rows`` (a BatchRows) restricts the last layer and the mean to those rows."""
        d = X0.shape[1]
        if out_mean is None:
            out_mean = torch.empty_like(X0)
        if cl_layer > 0 and out_cl is None:
            out_cl = torch.empty_like(X0)
        r = rows
        check(_lib.lib().idg_propagate_fwd_ex(self._h, ptr(X0), d, K, int(include_layer0), ptr(noise), float(eps), cl_layer,
                                              ptr(out_mean), ptr(out_cl), ptr(self.work(d)),
                                              ptr(r.rowlist) if r else None, ptr(r.count) if r else None, r.max_rows if r else 0,
                                              ptr(r.worklist(self)) if r else None, cur_stream()), "idg_propagate_fwd_ex")
        return (out_mean, out_cl) if cl_layer > 0 else out_mean

    def propagate_bwd(self, G, K, include_layer0, Gcl=None, cl_layer=0, out=None, rows=None, adam=None):
        """Backward w.r.t. X0.  ``rows``: G (and Gcl) are zero outside those rows -> sparse-input first product.
        ``adam`` (an _lib.AdamArgs): the last product applies the Adam update in its epilogue, no gradient is written."""
        d = G.shape[1]
        if adam is not None:
            import ctypes as _C
            check(_lib.lib().idg_propagate_bwd_adam(self._h, ptr(G), ptr(Gcl), d, K, int(include_layer0), cl_layer, ptr(self.work(d)),
                                                    ptr(rows.bitmap) if rows else None, _C.byref(adam), cur_stream()), "idg_propagate_bwd_adam")
            return None
        if out is None:
            out = torch.empty_like(G)
        check(_lib.lib().idg_propagate_bwd_ex(self._h, ptr(G), ptr(Gcl), d, K, int(include_layer0), cl_layer, ptr(out),
                                              ptr(self.work(d)), ptr(rows.bitmap) if rows else None, cur_stream()), "idg_propagate_bwd_ex")
        return out


def expected_closure_fraction(csr, num_users, batch):
    """Rough size of {batch rows + neighbours} / N for one mini-batch of `batch` (user, pos, neg) samples: users are
    drawn per train edge (degree-biased), positives are degree-biased items, negatives uniform items."""
    deg = (csr.indptr[1:] - csr.indptr[:-1]).double()
    du, di = deg[:num_users], deg[num_users:]
    biased = lambda x: float((x * x).sum() / x.sum().clamp(min=1.0))
    est = batch * (biased(du) + biased(di) + float(di.mean()))
    return min(1.0, est / float(deg.numel()))


class BatchRows:
    """Unique rows {user, U+pos, U+neg} of a mini-batch on the device: list + count + bitmap."""

    def __init__(self, N, max_batch, device):
        self.max_rows = 3 * max_batch
        self.rowlist = torch.zeros(self.max_rows, dtype=torch.int32, device=device)
        self.count = torch.zeros(1, dtype=torch.int32, device=device)
        self.bitmap = torch.zeros((N + 31) // 32 + 1, dtype=torch.int32, device=device)
        self.lead = torch.zeros(3 * max_batch, dtype=torch.uint8, device=device)
        self.closure = None  # optional bitmap: batch rows + their neighbours (enable_closure)
        self._wl = {}

    def worklist(self, graph, slot=0):
        """Schedule scratch of the row-restricted layer over ``graph``; launches that run concurrently (one stream each) take
        different ``slot``s -- the list is built with a counter, two launches must not share it."""
        k = (id(graph), slot)
        if k not in self._wl:
            n = int(_lib.lib().idg_graph_worklist_ints(graph._h, self.max_rows))
            self._wl[k] = torch.zeros(n, dtype=torch.int32, device=self.rowlist.device)
        return self._wl[k]

    def build(self, users_ptr, pos_ptr, neg_ptr, B, num_users):
        check(_lib.lib().idg_batch_rows(users_ptr, pos_ptr, neg_ptr, B, num_users, ptr(self.rowlist), ptr(self.count), ptr(self.bitmap),
                                        cur_stream()), "idg_batch_rows")

    def enable_closure(self, graph, buffer=None):
        """Allocate (or adopt) the closure bitmap and register it with the propagation handle."""
        self.closure = buffer if buffer is not None else torch.zeros_like(self.bitmap)
        check(_lib.lib().idg_graph_set_closure(graph._h, ptr(self.closure)), "idg_graph_set_closure")

    def build_closure(self, graph):
        if graph.is_whole:   # from the batch side: ~3 B neighbour lists instead of a pass over every nonzero
            check(_lib.lib().idg_closure_from_rows(graph._h, ptr(self.rowlist), ptr(self.count), self.max_rows, ptr(self.bitmap), ptr(self.closure),
                                                   cur_stream()), "idg_closure_from_rows")
        else:
            check(_lib.lib().idg_closure_bitmap(graph._h, ptr(self.bitmap), ptr(self.closure), cur_stream()), "idg_closure_bitmap")

    def build_unique(self, users_ptr, pos_ptr, neg_ptr, B, num_users, uidx, ucnt, iidx, icnt):
        check(_lib.lib().idg_batch_rows_unique(users_ptr, pos_ptr, neg_ptr, B, num_users, ptr(self.rowlist), ptr(self.count), ptr(self.bitmap),
                                               ptr(self.lead), ptr(uidx), ptr(ucnt), ptr(iidx), ptr(icnt), cur_stream()), "idg_batch_rows_unique")

    def clear(self):
        check(_lib.lib().idg_batch_rows_clear(ptr(self.rowlist), ptr(self.count), self.max_rows, ptr(self.bitmap), cur_stream()),
              "idg_batch_rows_clear")
        if self.closure is not None:
            self.closure.zero_()
