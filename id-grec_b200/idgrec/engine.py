"""Fused training step of the propagation models (LightGCN / SimGCL / XSimGCL / MFBPR).

One object owns the device state of a model's hot path -- the fused [N,d] parameter table
(user rows then item rows; ``user_embedding.weight`` / ``item_embedding.weight`` are views of
it), its gradient, the Adam moments, the layer work buffers -- and runs
    propagate -> BPR (+InfoNCE) -> backward propagate -> Adam
through the C ABI without autograd, Python-side tensor ops or host syncs, optionally replayed
from a CUDA graph.  It is what ``utility_train.trainer.universal_trainer`` drives; the autograd
path in the model classes computes the same values for the reference's own trainer loop
(trainer.py:40-56).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, cur_stream, ptr


# LightGCN-backbone models whose extra loss is a batch x batch term on (F_u[user], F_i[positive]) (SURVEY 8 f rank 4):
#   kind -> (loss slots reported, in the reference's order; BPR upstream weight; reg upstream weight; reg mask)
# raw slots of self.loss: 0 = bpr, 1 = reg, 2 / 3 = pair terms
PAIR_MODELS = {
    "LightCCF": ((0, 1, 2), 1.0, 1.0, 7),    # [bpr, reg, ssl_lambda * na]            models/LightCCF.py:73-77
    "LightCSCF": ((1, 2), 0.0, 1.0, 7),      # [reg, lambda_gamma * na]               models/LightCSCF.py:84-89 (LightGCN encoder)
    "SCCF": ((2, 3), 0.0, 0.0, 0),           # [-up, down]                            models/SCCF.py:80
    "DirectAU": ((2, 3, 1), 0.0, 1.0, 3),    # [align, gamma * uniform, reg(u, pos)]  models/DirectAU.py:69-78
}


class EpochFetch:
    """Shared by the fused trainers: recognise consecutive slices of the epoch's sample arrays so that the captured step can fetch
    its own mini-batch (idg_batch_fetch).  Needs self.d_step (device step counter, +1 per applied step), self.step_count (its host
    twin), self.max_batch, and self._ep / self._ep_ptrs (None / device int64[4])."""

    def _epoch_cursor(self, B, users, pos, neg):
        """trainer.py:40-47 walks the epoch's three (shuffled) sample arrays in slices of batch_size.  When the three batch tensors
        are such slices -- views at the same offset of three contiguous int64 arrays, taken in order from offset 0 -- the step graph
        fetches its batch itself (idg_batch_fetch, first node) and the host only replays.  Returns True if this call is the next
        slice of the registered epoch; registers a new epoch when it is the slice at offset 0."""
        bases = (users._base, pos._base, neg._base)
        if any(b is None for b in bases) or self.d_step is None:
            return False
        offs, starts = [], []
        for t, b in zip((users, pos, neg), bases):
            # the array is the base itself (1-D) or one row of a contiguous 2-D base (trainer.py's [3, E] sample block)
            if t.dtype != torch.int64 or b.dtype != torch.int64 or not t.is_contiguous() or not b.is_contiguous() or b.dim() not in (1, 2):
                return False
            cols = b.shape[-1]
            o = (t.data_ptr() - b.data_ptr()) // 8
            if o % cols + B > cols:
                return False
            offs.append(o % cols)
            starts.append(b.data_ptr() + (o // cols) * cols * 8)
        if offs[0] != offs[1] or offs[0] != offs[2]:
            return False
        key = tuple(starts)
        if offs[0] == 0 and B == min(self.max_batch, bases[0].shape[-1]):
            # first slice of an epoch: (re)register -- also when the same arrays are walked again
            self._ep = (key, self.step_count, B, bases)
            self._ep_ptrs.copy_(torch.tensor(list(key) + [self.step_count], dtype=torch.int64), non_blocking=False)
            return True
        ep = self._ep
        return ep is not None and ep[0] == key and B <= ep[2] and offs[0] == (self.step_count - ep[1]) * ep[2]

    def _fetch_batch(self, B, stride):
        check(self.l.idg_batch_fetch(ptr(self._ep_ptrs), ptr(self.d_step), stride, B, ptr(self.batch), self.batch.shape[1], cur_stream()), "idg_batch_fetch")


class FusedTrainer(EpochFetch):
    def __init__(self, kind, graph, table, num_users, K, reg_lambda, lr, ssl_lambda=0.0, temperature=0.2,
                 eps=0.0, cl_layer=1, max_batch=2048, use_cuda_graph=True, betas=(0.9, 0.999), adam_eps=1e-8,
                 restrict_rows=True, fuse_adam=True, closure_restrict="auto", margin=0.0, gamma=0.0):
        assert kind in ("LightGCN", "SimGCL", "XSimGCL", "MFBPR") or kind in PAIR_MODELS
        self.l = _lib.lib()
        self.kind, self.graph, self.E0 = kind, graph, table
        self.U, (self.N, self.d), self.K = num_users, table.shape, K
        self.reg_lambda, self.lr, self.betas, self.adam_eps = reg_lambda, lr, betas, adam_eps
        self.ssl_lambda, self.temperature, self.eps, self.cl_layer = ssl_lambda, temperature, eps, cl_layer
        dev = table.device
        self.dev = dev
        z = lambda: torch.zeros_like(table)
        self.gE0, self.m, self.v, self.G = z(), z(), z(), z()
        self.F = table if kind == "MFBPR" else torch.empty_like(table)
        self.step_count = 0
        self.max_batch = max_batch
        self.ws = torch.empty(int(self.l.idg_bpr_workspace_bytes(max_batch)), dtype=torch.uint8, device=dev)
        self.n_loss = 3 if kind in ("SimGCL", "XSimGCL") else 2
        self.loss_order = None
        if kind in PAIR_MODELS:
            order, w_bpr, w_reg, self.reg_mask = PAIR_MODELS[kind]
            self.loss_order = torch.tensor(order, dtype=torch.long, device=dev)
            self.n_loss = len(order)
            self.up_w = torch.tensor([w_bpr, w_reg], dtype=torch.float32, device=dev)
            self.margin, self.gamma = margin, gamma
            self.pA, self.pP = (torch.empty(max_batch, self.d, dtype=torch.float32, device=dev) for _ in range(2))
            self.gA, self.gP, self.gA2, self.gP2 = (torch.empty(max_batch, self.d, dtype=torch.float32, device=dev) for _ in range(4))
            self.pair_ws = torch.empty(int(self.l.idg_pair_loss_workspace_bytes(max_batch, self.d)), dtype=torch.uint8, device=dev)
            self.ucnt = torch.zeros(1, dtype=torch.int32, device=dev)
            self.icnt = torch.zeros(1, dtype=torch.int32, device=dev)
            self.uidx = torch.zeros(max_batch, dtype=torch.int64, device=dev)
            self.iidx = torch.zeros(max_batch, dtype=torch.int64, device=dev)
        self.loss = torch.zeros(4, dtype=torch.float32, device=dev)
        self.loss_acc = torch.zeros(4, dtype=torch.float64, device=dev)   # epoch sums in float64 (trainer.py:52-53)
        self.batch = torch.zeros(3, max_batch, dtype=torch.int64, device=dev)
        if kind in ("SimGCL", "XSimGCL"):
            self.noise = torch.empty(K, self.N, self.d, dtype=torch.float32, device=dev)
            self.noise_b = torch.empty(K, self.N, self.d, dtype=torch.float32, device=dev) if kind == "SimGCL" else None
            self.nce_ws = torch.empty(int(self.l.idg_infonce_workspace_bytes(max_batch, self.d)), dtype=torch.uint8, device=dev)
            # the user-side and item-side contrast terms touch disjoint gradient rows: second workspace + stream, run concurrently
            self.nce_ws2 = torch.empty_like(self.nce_ws)
            self._nce_side = torch.cuda.Stream(device=dev)
            self.V1 = torch.empty_like(table)
            self.V2 = torch.empty_like(table) if kind == "SimGCL" else None
            self.Gcl = z() if kind == "XSimGCL" else None
            # torch.unique(user) / torch.unique(positive) on the device, no host sync (SimGCL.py:80-81)
            self.uidx = torch.zeros(max_batch, dtype=torch.int64, device=dev)
            self.iidx = torch.zeros(max_batch, dtype=torch.int64, device=dev)
            self.ucnt = torch.zeros(1, dtype=torch.int32, device=dev)
            self.icnt = torch.zeros(1, dtype=torch.int32, device=dev)
        # identical-result work skipping (SURVEY.md 8 d): last forward layer only on the batch rows, first
        # backward product only over the batch columns
        self.rows = None
        self.use_closure = False
        if kind != "MFBPR":
            graph.work(self.d)  # allocate the layer ping-pong buffers before any graph capture
            if restrict_rows:
                from .graph import BatchRows
                self.rows = BatchRows(self.N, max_batch, dev)
                self.rows.worklist(graph)
                # layer K-1 / second backward product restricted to the batch neighbourhood when that is a small
                # part of the graph (scale-up graphs; ~76 % of the nodes at the amazon-book shape -> off)
                if closure_restrict == "auto":
                    from .graph import expected_closure_fraction
                    closure_restrict = K >= 2 and expected_closure_fraction(graph.csr, num_users, max_batch) < 0.4
                self.use_closure = bool(closure_restrict) and K >= 2
                if self.use_closure:
                    self.rows.enable_closure(graph)
        # Adam applied in the epilogue of the last backward layer (no gradient pass through memory)
        self.fuse_adam = fuse_adam and kind != "MFBPR"
        self.d_step = torch.zeros(1, dtype=torch.int32, device=dev)
        self._ep, self._ep_ptrs = None, torch.zeros(4, dtype=torch.int64, device=dev)     # registered epoch arrays (see _epoch_cursor)
        self.regc = torch.zeros(self.N, dtype=torch.float32, device=dev)
        self.adam_scalars = torch.zeros(2, dtype=torch.float32, device=dev)
        self.adam_args = _lib.AdamArgs(ptr(self.E0), ptr(self.m), ptr(self.v), ptr(self.regc), ptr(self.adam_scalars), betas[0], betas[1], adam_eps)
        # scalar tail of the BPR loss reduction: epoch loss sums (models whose loss list is complete after BPR) and the
        # bias-corrected Adam scalars of the step -- two single-thread launches less per step
        self._tail_acc = kind in ("LightGCN", "MFBPR")
        self._tail = {}
        for with_adam in (False, True):
            self._tail[with_adam] = _lib.StepTail(ptr(self.loss_acc) if self._tail_acc else None, ptr(self.d_step) if with_adam else None,
                                                  ptr(self.adam_scalars) if with_adam else None, lr, betas[0], betas[1])
        # the batch row set does not depend on the first K-1 propagation layers: it is built on a side stream (a parallel
        # branch of the captured graph) while they run
        self._side = torch.cuda.Stream(device=dev) if (self.rows is not None and kind != "MFBPR") else None
        # SimGCL: the batch-row last layers of the three propagations run concurrently, one stream each.  A propagation handle
        # owns the scratch of its heavy rows (chunk partials + arrival counters), so concurrent launches need a handle each.
        self._gviews = None
        if kind == "SimGCL" and self.rows is not None and not self.use_closure and 2 <= K <= 4:
            from .graph import Graph
            self._gviews = [graph, Graph(graph.csr, graph.row_begin, graph.row_end), Graph(graph.csr, graph.row_begin, graph.row_end)]
            for v, gv in enumerate(self._gviews):     # schedule scratch allocated (and zero-filled) here, not on first use: a lazy
                self.rows.worklist(gv, v)              # torch.zeros on the main stream would race with the launch on the view's stream
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        self._graph_launches = {}
        self.replayed_launches = 0  # kernels launched through graph replays (bench.py gpu_launches)
        self.injected_noise = None  # parity tests: list of per-view [K,N,d] tensors

    # ------------------------------------------------------------------ pieces
    def _adam(self):
        check(self.l.idg_adam_step_dev(ptr(self.E0), ptr(self.gE0), ptr(self.m), ptr(self.v), self.E0.numel(), self.lr,
                                       self.betas[0], self.betas[1], self.adam_eps, ptr(self.d_step), cur_stream()), "idg_adam_step_dev")

    def _bpr(self, B, u, p, n, fused):
        l, s = self.l, cur_stream()
        mask = getattr(self, "reg_mask", 7)
        up = ptr(self.up_w) if self.loss_order is not None else None     # which of (bpr, reg) enter the objective
        check(l.idg_bpr_forward_tail(ptr(self.F), ptr(self.E0), u, p, n, B, self.U, self.N, self.d, self.reg_lambda, mask,
                                     ptr(self.loss), C.byref(self._tail[bool(fused)]), ptr(self.ws), s), "idg_bpr_forward_tail")
        check(l.idg_bpr_backward(ptr(self.F), B, self.d, mask, up, ptr(self.G), self.reg_lambda, ptr(self.regc) if fused else None,
                                 ptr(self.ws), s), "idg_bpr_backward")

    def _finish(self, B, fused):
        # also clears the batch-row bitmap: the rows it visits are exactly the rows of this step's row set
        up = ptr(self.up_w) if self.loss_order is not None else None
        check(self.l.idg_bpr_finish_clear(ptr(self.E0), None if fused else ptr(self.gE0), ptr(self.G), B, self.d, self.reg_lambda, up,
                                          ptr(self.regc) if fused else None, ptr(self.rows.bitmap) if self.rows is not None else None,
                                          ptr(self.ws), cur_stream()), "idg_bpr_finish_clear")

    def _pair(self, B, u, p):
        """Batch x batch term(s) of the LightGCN-backbone models on (F[user], F[U + positive]): gather -> idg_pair_loss_ex
        (tcgen05 contractions) -> deterministic scatter-add of the row gradients into G, which the shared backward
        propagation then carries to the parameters together with the BPR gradient."""
        l, s, d = self.l, cur_stream(), self.d
        item_off = self.U * d * 4
        check(l.idg_gather_rows(ptr(self.F), u, B, d, ptr(self.pA), s), "idg_gather_rows")
        check(l.idg_gather_rows(ptr(self.F) + item_off, p, B, d, ptr(self.pP), s), "idg_gather_rows")
        self.loss[2:4].zero_()
        L2, L3 = ptr(self.loss[2:]), ptr(self.loss[3:])

        def pair(kind, X, Y, p0, p1, scale, out, gX, gY, ca=None, cb=None):
            check(l.idg_pair_loss_ex(kind, ptr(X), ptr(Y) if Y is not None else None, B, d, p0, p1, scale, ca, cb, out, ptr(gX),
                                     ptr(gY) if gY is not None else None, ptr(self.pair_ws), s), "idg_pair_loss_ex")

        def scatter(gX, idx, off):
            check(l.idg_scatter_add_rows(ptr(gX), idx, B, d, ptr(self.G) + off, s), "idg_scatter_add_rows")

        if self.kind == "LightCCF":
            pair(0, self.pA, self.pP, self.temperature, 0.0, self.ssl_lambda, L2, self.gA, self.gP)
        elif self.kind == "LightCSCF":
            pair(1, self.pA, self.pP, self.temperature, self.margin, self.ssl_lambda, L2, self.gA, self.gP)
        elif self.kind == "SCCF":
            pair(3, self.pA, self.pP, self.temperature, 0.0, 1.0, L2, self.gA, self.gP)
            pair(2, self.pA, self.pP, self.temperature, 0.0, 1.0, L3, self.gA2, self.gP2, ptr(self.ucnt), ptr(self.icnt))
            check(l.idg_axpby(ptr(self.gA), 1.0, ptr(self.gA), 1.0, ptr(self.gA2), B * d, s), "idg_axpby")
            check(l.idg_axpby(ptr(self.gP), 1.0, ptr(self.gP), 1.0, ptr(self.gP2), B * d, s), "idg_axpby")
        else:  # DirectAU: align + gamma * (uniform(users) + uniform(items)) / 2
            pair(4, self.pA, self.pP, 0.0, 0.0, 1.0, L2, self.gA, self.gP)
            pair(5, self.pA, None, 0.0, 0.0, 0.5 * self.gamma, L3, self.gA2, None)
            pair(5, self.pP, None, 0.0, 0.0, 0.5 * self.gamma, L3, self.gP2, None)
            check(l.idg_axpby(ptr(self.gA), 1.0, ptr(self.gA), 1.0, ptr(self.gA2), B * d, s), "idg_axpby")
            check(l.idg_axpby(ptr(self.gP), 1.0, ptr(self.gP), 1.0, ptr(self.gP2), B * d, s), "idg_axpby")
        scatter(self.gA, u, 0)
        scatter(self.gP, p, item_off)

    def _draw_noise(self, view, buf=None):
        buf = self.noise if buf is None else buf
        if self.injected_noise is not None:
            buf.copy_(self.injected_noise[view])
        else:  # same call pattern as SimGCL.py:50: one rand_like([N,d]) per layer, views in order
            for k in range(self.K):
                buf[k].uniform_()

    def _build_rows(self, B, u, p, n, contrastive):
        rows = self.rows
        if contrastive:
            rows.build_unique(u, p, n, B, self.U, self.uidx, self.ucnt, self.iidx, self.icnt)
        else:
            rows.build(u, p, n, B, self.U)

    def _forward_mean_on_batch_rows(self):
        """LightGCN.py:36-52 for a step whose loss reads the final mean at the batch rows only: layers 1..K-1 are plain
        products (no running-sum traffic: the full-size layer sum is never formed), the last layer and the mean
        ((E0 + X1) + X2 + X3) / (K+1) -- the reference's summation order -- are evaluated on the batch rows."""
        g, K, d, rows = self.graph, self.K, self.d, self.rows
        w = g.work(d)
        bufs = [w[:self.N * d].view(self.N, d), w[self.N * d:].view(self.N, d)]
        x, layers = self.E0, [self.E0]
        for k in range(K - 1):
            y = bufs[k & 1]
            g.spmm_layer(x, Y=y)
            layers.append(y)
            x = y
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)     # join: the row set is ready
        layers += [None] * (3 - len(layers))
        check(self.l.idg_spmm_layer_rows(g._h, ptr(x), None, None, 0.0, ptr(layers[0]), ptr(layers[1]), ptr(layers[2]), ptr(self.F), float(K + 1), d,
                                         ptr(rows.rowlist), ptr(rows.count), rows.max_rows, ptr(rows.worklist(g)), cur_stream()), "idg_spmm_layer_rows")

    def _simgcl_forward_shared(self):
        """SimGCL.py:62-66 runs aggregate() three times on the same ego table (clean, view 1, view 2).  Their first product
        A_hat . E0 is the same gather: one launch writes the clean row and both perturbed copies (idg_spmm_layer_views); the
        later layers run per propagation; the last layer and the mean over layers 1..K (SimGCL.py:45,55-56: no layer 0) are
        evaluated on the batch rows.  Bit-identical to three separate propagations."""
        g, K, d, N, rows, l = self.graph, self.K, self.d, self.N, self.rows, self.l
        if getattr(self, "_sg", None) is None:
            self._sg = [[torch.empty(N, d, dtype=torch.float32, device=self.dev) for _ in range(K - 1)] for _ in range(3)]
        X = self._sg
        nz = (None, self.noise, self.noise_b)
        main, s1, s2 = torch.cuda.current_stream(), self._side, self._nce_side
        # Noise draws in the reference's order (view 1 layers 0..K-1, then view 2: the generator is advanced on the host, call by
        # call), but only the two first-layer draws sit in front of the first product: the others fill their buffers from a second
        # stream while that product runs.
        if self.injected_noise is not None:
            self._draw_noise(0)
            self._draw_noise(1, self.noise_b)
        else:
            s2.wait_stream(main)
            for buf in (self.noise, self.noise_b):
                buf[0].uniform_()
                with torch.cuda.stream(s2):
                    for k in range(1, K):
                        buf[k].uniform_()
        check(l.idg_spmm_layer_views(g._h, ptr(self.E0), ptr(X[0][0]), ptr(self.noise[0]), ptr(X[1][0]), ptr(self.noise_b[0]), ptr(X[2][0]),
                                     float(self.eps), d, cur_stream()), "idg_spmm_layer_views")
        if self.injected_noise is None:
            main.wait_stream(s2)                                    # join: the later layers' noise is drawn
        for k in range(1, K - 1):
            for v in range(3):
                g.spmm_layer(X[v][k - 1], Y=X[v][k], noise=None if v == 0 else nz[v][k], eps=0.0 if v == 0 else self.eps)
        main.wait_stream(s1)                                        # join: the row set is ready
        # the three batch-row last layers are latency-bound (a few thousand rows each): one per stream
        s1.wait_stream(main)
        s2.wait_stream(main)
        for v, (out, st) in enumerate(((self.F, main), (self.V1, s1), (self.V2, s2))):
            acc = list(X[v]) + [None] * (3 - (K - 1))
            g = self._gviews[v]
            check(l.idg_spmm_layer_rows(g._h, ptr(X[v][K - 2]), None, None if v == 0 else ptr(nz[v][K - 1]), 0.0 if v == 0 else float(self.eps),
                                        ptr(acc[0]), ptr(acc[1]), ptr(acc[2]), ptr(out), float(K), d, ptr(rows.rowlist), ptr(rows.count), rows.max_rows,
                                        ptr(rows.worklist(g, v)), st.cuda_stream), "idg_spmm_layer_rows")
        main.wait_stream(s1)
        main.wait_stream(s2)

    def _xsimgcl_forward_split(self):
        """XSimGCL.py:69-82 for a step whose losses read the final mean (and, for cl_layer = K, the contrast view) at the batch
        rows only: layers 1..K-1 are plain perturbed products, the last layer and the mean over layers 1..K run on the batch rows
        (same launches as the row-partitioned trainer, dist.py:_forward).  Only the first layer's noise draw sits in front of the
        first product; the later layers' draws fill their buffers from a second stream meanwhile (draw ORDER on the host, i.e.
        the generator stream, is the reference's).  Returns the contrast view (post-noise output of layer cl_layer)."""
        g, K, d, N, rows, l = self.graph, self.K, self.d, self.N, self.rows, self.l
        if getattr(self, "_xg", None) is None:
            self._xg = [torch.empty(N, d, dtype=torch.float32, device=self.dev) for _ in range(K - 1)]
        W = self._xg
        main, s2 = torch.cuda.current_stream(), self._nce_side
        overlap = self.injected_noise is None
        if overlap:
            s2.wait_stream(main)
            self.noise[0].uniform_()
            with torch.cuda.stream(s2):
                for k in range(1, K):
                    self.noise[k].uniform_()
        else:
            self._draw_noise(0)
        x = self.E0
        for k in range(K - 1):
            g.spmm_layer(x, Y=W[k], noise=self.noise[k], eps=self.eps)
            if k == 0 and overlap:
                main.wait_stream(s2)
            x = W[k]
        main.wait_stream(self._side)                                # join: the row set is ready
        acc = list(W) + [None] * (3 - (K - 1))
        y_cl = self.V1 if self.cl_layer == K else None
        check(l.idg_spmm_layer_rows(g._h, ptr(x), ptr(y_cl), ptr(self.noise[K - 1]), float(self.eps), ptr(acc[0]), ptr(acc[1]), ptr(acc[2]), ptr(self.F),
                                    float(K), d, ptr(rows.rowlist), ptr(rows.count), rows.max_rows, ptr(rows.worklist(g)), cur_stream()), "idg_spmm_layer_rows")
        return self.V1 if self.cl_layer == K else W[self.cl_layer - 1]

    def _contrast_pair(self, uniq, B, Va, Vb, gA, gB):
        """InfoNCE over the batch's unique users and over its unique positives (SimGCL.py:80-88, XSimGCL.py:84-92).  The two
        terms read the same views and accumulate into disjoint rows (users < U <= items), so the item term runs on a second
        stream with its own workspace; both add into the (zeroed) loss slot atomically -- two addends, order-free."""
        l, main = self.l, torch.cuda.current_stream()
        (uidx, ucnt), (iidx, icnt) = uniq
        self._nce_side.wait_stream(main)
        with torch.cuda.stream(self._nce_side):
            check(l.idg_infonce_fwd_bwd_dev(ptr(Va), ptr(Vb), ptr(iidx), ptr(icnt), B, self.d, self.temperature, self.ssl_lambda,
                                            ptr(self.loss[2:]), ptr(gA), ptr(gB), ptr(self.nce_ws2), cur_stream()), "idg_infonce_fwd_bwd_dev")
        check(l.idg_infonce_fwd_bwd_dev(ptr(Va), ptr(Vb), ptr(uidx), ptr(ucnt), B, self.d, self.temperature, self.ssl_lambda,
                                        ptr(self.loss[2:]), ptr(gA), ptr(gB), ptr(self.nce_ws), cur_stream()), "idg_infonce_fwd_bwd_dev")
        main.wait_stream(self._nce_side)

    def _body(self, B, u, p, n, users_t=None, pos_t=None, fused=False):
        """Kernels of one step for batch pointers u/p/n (device int64).  ``fused``: Adam inside the last backward layer."""
        g, K = self.graph, self.K
        adam = self.adam_args if fused else None
        out = None if fused else self.gE0
        rows = self.rows
        contrastive = self.kind in ("SimGCL", "XSimGCL") or self.kind == "SCCF"   # SCCF needs the two unique counts
        # forward of the LightGCN-encoder steps without the full-size layer sum, row set built on a parallel branch
        split_fwd = (rows is not None and not self.use_closure and 2 <= K <= 3 and (self.kind == "LightGCN" or self.kind in PAIR_MODELS))
        shared_fwd = rows is not None and not self.use_closure and 2 <= K <= 4 and self.kind == "SimGCL"
        xsplit = rows is not None and not self.use_closure and 2 <= K <= 4 and self.kind == "XSimGCL"
        split_fwd = split_fwd or shared_fwd or xsplit
        if rows is not None and split_fwd and self._side is not None:
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)                            # fork: after the batch copy / the previous step's clean-up
            with torch.cuda.stream(self._side):
                self._build_rows(B, u, p, n, contrastive)
        elif rows is not None:
            self._build_rows(B, u, p, n, contrastive)
        if self.use_closure:
            rows.build_closure(g)
        if self.kind == "LightGCN":
            if split_fwd:
                self._forward_mean_on_batch_rows()
            else:
                g.propagate_fwd(self.E0, K, True, out_mean=self.F, rows=rows)
            self._bpr(B, u, p, n, fused)
            g.propagate_bwd(self.G, K, True, out=out, rows=rows, adam=adam)
        elif self.kind in PAIR_MODELS:
            if rows is None and self.kind == "SCCF":
                check(self.l.idg_unique_rows(u, B, 0, ptr(self.uidx), ptr(self.ucnt), cur_stream()), "idg_unique_rows")
                check(self.l.idg_unique_rows(p, B, self.U, ptr(self.iidx), ptr(self.icnt), cur_stream()), "idg_unique_rows")
            if split_fwd:
                self._forward_mean_on_batch_rows()
            else:
                g.propagate_fwd(self.E0, K, True, out_mean=self.F, rows=rows)
            self._bpr(B, u, p, n, fused)
            self._pair(B, u, p)
            g.propagate_bwd(self.G, K, True, out=out, rows=rows, adam=adam)
        elif self.kind == "MFBPR":
            self._bpr(B, u, p, n, False)
            self.gE0.copy_(self.G)
        else:
            l, s = self.l, cur_stream()
            if rows is None:  # no batch row set: stand-alone (sorted) unique
                check(l.idg_unique_rows(u, B, 0, ptr(self.uidx), ptr(self.ucnt), s), "idg_unique_rows")
                check(l.idg_unique_rows(p, B, self.U, ptr(self.iidx), ptr(self.icnt), s), "idg_unique_rows")
            uniq = ((self.uidx, self.ucnt), (self.iidx, self.icnt))
            if self.kind == "SimGCL":
                if shared_fwd:
                    self._simgcl_forward_shared()
                else:
                    g.propagate_fwd(self.E0, K, False, out_mean=self.F, rows=rows)
                    self._draw_noise(0)
                    g.propagate_fwd(self.E0, K, False, noise=self.noise, eps=self.eps, out_mean=self.V1, rows=rows)
                    self._draw_noise(1)
                    g.propagate_fwd(self.E0, K, False, noise=self.noise, eps=self.eps, out_mean=self.V2, rows=rows)
                self._bpr(B, u, p, n, fused)
                self.loss[2:3].zero_()
                # the three propagations share one linear backward operator: accumulate all row
                # gradients into G and back-propagate once (9 backward SpMMs of the reference -> 3)
                self._contrast_pair(uniq, B, self.V1, self.V2, self.G, self.G)
                g.propagate_bwd(self.G, K, False, out=out, rows=rows, adam=adam)
            else:  # XSimGCL: one perturbed propagation, contrast view captured at cl_layer
                if xsplit:
                    cl_view = self._xsimgcl_forward_split()
                else:
                    self._draw_noise(0)
                    g.propagate_fwd(self.E0, K, False, noise=self.noise, eps=self.eps, cl_layer=self.cl_layer, out_mean=self.F, out_cl=self.V1, rows=rows)
                    cl_view = self.V1
                self._bpr(B, u, p, n, fused)
                self.loss[2:3].zero_()
                self._contrast_pair(uniq, B, cl_view, self.F, self.Gcl, self.G)
                g.propagate_bwd(self.G, K, False, Gcl=self.Gcl, cl_layer=self.cl_layer, out=out, rows=rows, adam=adam)
                for idx, _ in uniq:  # entries past the count are stale but valid rows of an all-zero table: harmless
                    check(l.idg_zero_rows(ptr(self.Gcl), ptr(idx), B, self.d, s), "idg_zero_rows")
        self._finish(B, fused)
        if rows is not None and rows.closure is not None:
            rows.closure.zero_()

    # ------------------------------------------------------------------ public
    def step(self, users, pos, neg, apply_adam=True):
        """One training step on device int64 index tensors; losses land in self.loss[:n_loss]."""
        B = int(users.numel())
        assert B <= self.max_batch
        if self.use_cuda_graph and apply_adam:
            return self._step_graph(B, users, pos, neg)
        users, pos, neg = users.contiguous(), pos.contiguous(), neg.contiguous()
        fused = apply_adam and self.fuse_adam
        self._body(B, ptr(users), ptr(pos), ptr(neg), users, pos, fused=fused)
        if apply_adam:
            if not fused:
                self._adam()
            self.step_count += 1
        if not self._tail_acc:
            self.loss_acc += self.loss
        return self._report(self.loss)

    def _step_graph(self, B, users, pos, neg):
        """CUDA-graph replay: the step's kernels (incl. the device-side Adam step counter) are captured once per batch size.  Batches
        that are consecutive slices of the epoch's sample arrays are fetched by the graph itself; any other batch is copied into
        the static slab first."""
        fetch = self._epoch_cursor(B, users, pos, neg)
        key = ("f", B, self._ep[2]) if fetch else B
        if not fetch:
            self.batch[0, :B].copy_(users); self.batch[1, :B].copy_(pos); self.batch[2, :B].copy_(neg)
        if key not in self._graphs:
            self._capture(B, key)
        self._graphs[key].replay()
        self.replayed_launches += self._graph_launches[key]
        self.step_count += 1
        return self._report(self.loss)

    def _report(self, raw):
        """Losses in the order of the model's loss list (raw slots: bpr, reg, pair terms)."""
        return raw[:self.n_loss] if self.loss_order is None else raw[self.loss_order]

    def _capture(self, B, key=None):
        key = B if key is None else key
        u, p, n = (self.batch[k].data_ptr() for k in range(3))

        def run():
            if isinstance(key, tuple):      # ("f", B, stride): the batch comes from the registered epoch arrays
                self._fetch_batch(B, key[2])
            self._body(B, u, p, n, fused=self.fuse_adam)
            if not self.fuse_adam:
                self._adam()
            if not self._tail_acc:
                check(self.l.idg_accumulate_f64(ptr(self.loss_acc), ptr(self.loss), 4, cur_stream()), "idg_accumulate_f64")

        # warm-up outside capture would advance the model; capture directly (kernels are launched lazily at replay)
        torch.cuda.synchronize()
        n0 = self.l.idg_launch_count()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            run()
        self._graph_launches[key] = int(self.l.idg_launch_count() - n0)
        self._graphs[key] = gr

    def pop_epoch_losses(self):
        out = self._report(self.loss_acc).tolist()
        self.loss_acc.zero_()
        return out
