"""Fused training step of NGCF (models/NGCF.py:67-130 + trainer.py:40-56) without autograd or per-step allocations:

    3 x (SpMM -> fused dense layer kernel writing its 64-column block of the [N,256] concat in place)
    -> BPR on the 256-d rows + item-ego L2 -> 3 x (dense backward kernel -> SpMM with the direct term as addend)
    -> Adam on the fused [N,64] table and on ONE flat buffer holding the 12 dense weight tensors

captured once per batch size in a CUDA graph.  What the autograd path pays per step on top of the same kernels --
torch.cat of the layer outputs and its backward slices made contiguous, a fresh zero-filled [N,256] gradient, a
multi-tensor Adam with four passes over each table -- is gone; the arithmetic per element is unchanged
(tests/test_gpu_znext.py::test_ngcf_fused_step_equals_eager_loop)."""
from __future__ import annotations

import ctypes as C

import os

import torch

from . import _lib
from ._lib import check, cur_stream, ptr


class NgcfFusedTrainer:
    def __init__(self, model, lr, max_batch, use_cuda_graph=True, betas=(0.9, 0.999), adam_eps=1e-8):
        self.l = _lib.lib()
        self.model, self.graph, self.lr, self.betas, self.adam_eps = model, model.Graph, lr, betas, adam_eps
        self.max_batch, self.use_cuda_graph = max_batch, use_cuda_graph
        if model._table is None:
            model._fuse_tables()
        self.E0 = model._table
        if self.E0 is None or not self.E0.is_cuda:
            raise RuntimeError("model must be moved to a CUDA device first (model.to(device)); there is no CPU fallback")
        dev = self.E0.device
        self.dev, self.U, (self.N, self.d) = dev, model.dataset.num_users, self.E0.shape
        assert self.d == 64, "the fused NGCF step is written for embedding_size = layer_size = 64"
        self.K = int(model.config['GCN_layer'])
        self.reg_lambda = model.reg_lambda
        self.drop = [float(p) for p in getattr(model, "mess_dropout", [0.0] * self.K)]
        # ---- the 4 K dense tensors live in one flat buffer (parameters become views of it): one Adam launch for all of them
        names = []
        for layer in range(self.K):
            names += ['W_gcn_%d' % layer, 'b_gcn_%d' % layer, 'W_bi_%d' % layer, 'b_bi_%d' % layer]
        params = [model.weight_dict[k] for k in names]
        sizes = [p.numel() for p in params]
        pad = (-sum(sizes)) % 4
        self.flat = torch.zeros(sum(sizes) + pad, dtype=torch.float32, device=dev)
        self.w, self.gw, off = {}, {}, 0
        self.gflat = torch.zeros_like(self.flat)
        for k, p, n in zip(names, params, sizes):
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)
            self.w[k], self.gw[k] = p.data, self.gflat[off:off + n]
            off += n
        z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
        N, d, K = self.N, self.d, self.K
        self.final = z(N, (K + 1) * d)
        self.side, self.S, self.D = ([z(N, d) for _ in range(K)] for _ in range(3))
        if os.environ.get("IDG_NGCF_BWD") != "fma":
            self.S = None                                # the tensor-core backward does not read the pre-activations
        self.keep_all = z(K, N, d)                       # the K dropout masks of a step, drawn in one launch
        self.keep = [self.keep_all[k] for k in range(K)]
        self._keep_prob = (C.c_float * K)(*[1.0 - float(p) for p in self.drop[:K]])
        # the draws packed 64 bits per row for the tensor-core dense kernels (8 B per row instead of a 256 B float mask row, read in
        # the forward and again in the backward); the float masks remain for injected masks (parity tests) and IDG_NGCF_BWD=fma
        self.keep_bits = torch.zeros(K, N, 2, dtype=torch.int32, device=dev) if os.environ.get("IDG_NGCF_BWD") != "fma" else None
        self._seed = int(torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()].initial_seed()) & ((1 << 63) - 1)
        self.G = z(N, (K + 1) * d)                       # dL/dfinal: non-zero on the batch rows only, re-zeroed by bpr_finish
        self.G64 = z(N, d)                               # scratch for the reg-only BPR call (stays zero)
        self.dside, self.dEd = z(N, d), z(N, d)
        self.gE = [z(N, d), z(N, d)]                     # gradient w.r.t. a layer's input, ping-pong
        self.gE0 = z(N, d)
        self.m, self.v, self.mw, self.vw = z(N, d), z(N, d), torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.step_a = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_b = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ws256 = torch.empty(int(self.l.idg_bpr_workspace_bytes(max_batch)), dtype=torch.uint8, device=dev)
        self.ws64 = torch.empty(int(self.l.idg_bpr_workspace_bytes(max_batch)), dtype=torch.uint8, device=dev)
        self.ngws = torch.empty(int(self.l.idg_ngcf_workspace_bytes()), dtype=torch.uint8, device=dev)
        self.up_reg = torch.tensor([0.0, 1.0], dtype=torch.float32, device=dev)   # the d = 64 BPR call contributes the reg term only
        self.loss_a, self.loss_b = z(4), z(4)
        self.loss_acc = torch.zeros(2, dtype=torch.float64, device=dev)   # epoch sums in float64 (trainer.py:52-53)
        self.batch = torch.zeros(3, max_batch, dtype=torch.int64, device=dev)
        self.graph.work(d)
        self._graphs, self.replays, self.step_count = {}, 0, 0
        self.injected_keep = None     # parity tests: K [N,d] 0/1 masks

    # ------------------------------------------------------------------------------------------------------------
    def _S(self, layer):
        """Pre-activation buffer of a layer: only the CUDA-core backward (IDG_NGCF_BWD=fma) reads it; otherwise it is not even written."""
        return ptr(self.S[layer]) if self.S is not None else None

    def _body(self, B, u, p, n):
        l, s, g, K, d, N = self.l, cur_stream(), self.graph, self.K, self.d, self.N
        wd, W = self.w, (K + 1) * d
        fbase, gbase = ptr(self.final), ptr(self.G)
        # block 0 of the concat = the ego table, read by the BPR kernels at the batch rows only
        for idx, off in ((u, 0), (p, self.U), (n, self.U)):
            check(l.idg_copy_rows_strided(ptr(self.E0), d, idx, B, off, d, fbase, W, s), "idg_copy_rows_strided")
        bits = self.keep_bits is not None and self.injected_keep is None
        if bits:
            # nn.Dropout(p)'s draws (NGCF.py:99-100: always active): Bernoulli(1 - p) per element, all layers in one launch
            check(l.idg_ngcf_keep_bits(ptr(self.keep_bits), N, K, self._keep_prob, self._seed, ptr(self.step_a), s), "idg_ngcf_keep_bits")
        elif self.injected_keep is None:
            check(l.idg_ngcf_keep_masks(ptr(self.keep_all), N * d, K, self._keep_prob, self._seed, ptr(self.step_a), s), "idg_ngcf_keep_masks")
        E = self.E0
        for layer in range(K):
            pr = self.drop[layer]
            if self.injected_keep is not None:
                self.keep[layer].copy_(self.injected_keep[layer])
            g.spmm_layer(E, Y=self.side[layer])
            check((l.idg_ngcf_dense_fwd_bits if bits else l.idg_ngcf_dense_fwd)(
                ptr(E), ptr(self.side[layer]), ptr(wd['W_gcn_%d' % layer]), ptr(wd['b_gcn_%d' % layer]), ptr(wd['W_bi_%d' % layer]), ptr(wd['b_bi_%d' % layer]),
                ptr(self.keep_bits[layer]) if bits else ptr(self.keep[layer]), pr, N, self._S(layer), ptr(self.D[layer]), fbase + 4 * d * (layer + 1), W, s),
                "idg_ngcf_dense_fwd")
            E = self.D[layer]
        # BPR on the 256-d rows (NGCF.py:113-118); L2 on the item ego rows only (NGCF.py:120-125: reg mask 6)
        check(l.idg_bpr_forward(fbase, fbase, u, p, n, B, self.U, N, W, 0.0, 0, ptr(self.loss_a), ptr(self.ws256), s), "idg_bpr_forward")
        check(l.idg_bpr_forward(ptr(self.E0), ptr(self.E0), u, p, n, B, self.U, N, d, self.reg_lambda, 6, ptr(self.loss_b), ptr(self.ws64), s),
              "idg_bpr_forward")
        check(l.idg_bpr_backward(fbase, B, W, 0, None, gbase, 0.0, None, ptr(self.ws256), s), "idg_bpr_backward")
        check(l.idg_bpr_backward(ptr(self.E0), B, d, 6, ptr(self.up_reg), ptr(self.G64), self.reg_lambda, None, ptr(self.ws64), s), "idg_bpr_backward")
        ext = None
        for layer in range(K - 1, -1, -1):
            Ein = self.E0 if layer == 0 else self.D[layer - 1]
            out = self.gE[layer & 1]
            db = self.gw['b_gcn_%d' % layer]
            tail = (ptr(self.D[layer]), gbase + 4 * d * (layer + 1), W, ptr(ext) if ext is not None else None, N, ptr(self.dside), ptr(self.dEd),
                    ptr(self.gw['W_gcn_%d' % layer]), ptr(self.gw['W_bi_%d' % layer]), ptr(db), ptr(self.ngws), s)
            head = (ptr(Ein), ptr(self.side[layer]), ptr(wd['W_gcn_%d' % layer]), ptr(wd['W_bi_%d' % layer]))
            if bits:
                check(l.idg_ngcf_dense_bwd_bits(*head, ptr(self.keep_bits[layer]), self.drop[layer], *tail), "idg_ngcf_dense_bwd_bits")
            else:
                check(l.idg_ngcf_dense_bwd(*head, ptr(self.keep[layer]), self.drop[layer], self._S(layer), *tail), "idg_ngcf_dense_bwd")
            self.gw['b_bi_%d' % layer].copy_(db)                       # both biases enter the same sum (NGCF.py:91-95)
            g.spmm_layer(self.dside, Y=out, addend=self.dEd)           # d/dE_in = direct term + A_hat . dside (A_hat symmetric)
            ext = out
        # gradient of the ego table: through the layers + its own block of the concat + the L2 term
        torch.add(ext, self.G[:, :d], out=self.gE0)
        check(l.idg_bpr_finish(ptr(self.E0), ptr(self.gE0), ptr(self.G64), B, d, self.reg_lambda, ptr(self.up_reg), None, ptr(self.ws64), s), "idg_bpr_finish")
        check(l.idg_bpr_finish(None, None, gbase, B, W, 0.0, None, None, ptr(self.ws256), s), "idg_bpr_finish")
        self.loss_acc[0:1].add_(self.loss_a[0:1])
        self.loss_acc[1:2].add_(self.loss_b[1:2])

    def _adam(self):
        l, s = self.l, cur_stream()
        check(l.idg_adam_step_dev(ptr(self.E0), ptr(self.gE0), ptr(self.m), ptr(self.v), self.E0.numel(), self.lr, self.betas[0], self.betas[1],
                                  self.adam_eps, ptr(self.step_a), s), "idg_adam_step_dev")
        check(l.idg_adam_step_dev(ptr(self.flat), ptr(self.gflat), ptr(self.mw), ptr(self.vw), self.flat.numel(), self.lr, self.betas[0], self.betas[1],
                                  self.adam_eps, ptr(self.step_b), s), "idg_adam_step_dev")

    # ------------------------------------------------------------------------------------------------------------
    def step(self, users, pos, neg, apply_adam=True):
        """One training step on device int64 index tensors -> tensor [bpr, reg]."""
        B = int(users.numel())
        assert B <= self.max_batch
        self.batch[0, :B].copy_(users); self.batch[1, :B].copy_(pos); self.batch[2, :B].copy_(neg)
        u, p, n = (self.batch[k].data_ptr() for k in range(3))
        if self.use_cuda_graph and apply_adam and self.injected_keep is None:
            if B not in self._graphs:
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):   # captured directly: a warm-up pass would advance the model
                    self._body(B, u, p, n)
                    self._adam()
                self._graphs[B] = gr
            self._graphs[B].replay()
            self.replays += 1
        else:
            self._body(B, u, p, n)
            if apply_adam:
                self._adam()
        self.step_count += 1
        return torch.stack([self.loss_a[0], self.loss_b[1]])

    def pop_epoch_losses(self):
        out = self.loss_acc.tolist()
        self.loss_acc.zero_()
        return out
