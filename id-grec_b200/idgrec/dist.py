"""Row-partitioned multi-GPU training (SURVEY.md section 8 e): one process per GPU on one NVSwitch box.

Nodes (rows of the normalised adjacency, of every layer buffer, of the parameter table and of the
Adam state) are split into contiguous, nnz-balanced ranges.  Each propagation layer computes the
local rows and its SpMM epilogue stores every finished row into all peers' layer buffers over
NVLink (CUDA-IPC symmetric slab, csrc/peers.cu + SpmmArgs::peerY), followed by a device-side flag
barrier -- the per-layer all-gather is fused into the kernel that produces the rows.  The loss is
evaluated redundantly on every rank from the (complete) final rows of the mini-batch, so there is no
gradient reduction; each rank applies Adam to the rows it owns and pushes them to its peers.
torch.distributed (NCCL) is used only for bootstrap and scalar metric reductions.  Results are
bit-identical to one GPU: a row's reduction order is a function of the row alone.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import check, cur_stream, ptr
from .engine import EpochFetch
from .graph import BatchRows, Graph


def partition_rows(indptr, world: int):
    """Contiguous row ranges balanced by nonzeros (+ one unit per row for the epilogue traffic).
    Returns world+1 boundaries; inner boundaries are multiples of 128 rows (16-byte aligned row ranges, and each
    rank's slice of a row bitmap is a whole number of 16-byte groups of words)."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = len(indptr) - 1
    cost = indptr + 8 * np.arange(n + 1, dtype=np.int64)      # cumulative work up to each row boundary
    bounds = [0]
    for r in range(1, world):
        b = int(np.searchsorted(cost, cost[-1] * r / world))
        b = min(n, max(bounds[-1], (b + 127) // 128 * 128))
        bounds.append(b)
    bounds.append(n)
    return bounds


def shard_range(n: int, rank: int, world: int):
    """Contiguous shard of n units for user-sharded evaluation."""
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per)


class _CudaBuffer:
    """Exposes raw device memory to torch through __cuda_array_interface__."""

    def __init__(self, p, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(p), False), "version": 3, "strides": None}


class PeerSlab:
    """One slab per rank, mapped on every rank; tensors carved at identical offsets.

    Backends: ``symm`` -- torch.distributed._symmetric_memory (CUDA VMM); when the fabric supports it this also
    yields an NVSwitch multicast address, so a finished row leaves the GPU once (multimem.st) and the switch
    replicates it (NVLS).  ``ipc`` -- plain cudaMalloc + CUDA IPC handles, unicast peer stores.  IDG_SLAB
    selects one; the default tries ``symm`` and falls back to ``ipc``."""

    def __init__(self, nbytes: int, rank: int, world: int, group=None, device=None):
        import os
        import torch.distributed as dist
        self.l = _lib.lib()
        self.rank, self.world, self.nbytes = rank, world, int((nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20))
        self.device = device
        self.multicast = False
        mode = os.environ.get("IDG_SLAB", "auto")
        self._bytes = None
        if world > 1 and mode in ("auto", "symm"):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                t = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=device)
                hdl = symm_mem.rendezvous(t, dist.group.WORLD if group is None else group)
                t.zero_()
                self._symm = (t, hdl)
                self._bytes, self.base = t, t.data_ptr()
                self.peer_bases = [int(p) for p in hdl.buffer_ptrs]
                self.peer_bases[rank] = self.base
                mc = int(hdl.multicast_ptr) if os.environ.get("IDG_MULTICAST", "1") != "0" else 0
                self.backend = "symm"
            except Exception as e:  # noqa: BLE001 -- any failure of the optional backend falls back to IPC
                if mode == "symm":
                    raise
                self._bytes = None
                self._symm_error = repr(e)
        if self._bytes is None:
            mc = 0
            self.backend = "ipc"
            base = C.c_void_p()
            check(self.l.idg_device_alloc(self.nbytes, C.byref(base)), "idg_device_alloc")
            self.base = base.value
            handle = (C.c_char * 64)()
            check(self.l.idg_ipc_get_handle(self.base, handle), "idg_ipc_get_handle")
            handles = [None] * world
            if world > 1:
                dist.all_gather_object(handles, bytes(handle.raw), group=group)
            else:
                handles[0] = bytes(handle.raw)
            self.peer_bases = []
            for r in range(world):
                if r == rank:
                    self.peer_bases.append(self.base)
                else:
                    out = C.c_void_p()
                    buf = C.create_string_buffer(handles[r], 64)
                    check(self.l.idg_ipc_open(buf, C.byref(out)), "idg_ipc_open")
                    self.peer_bases.append(out.value)
            self._bytes = torch.as_tensor(_CudaBuffer(self.base, self.nbytes), device=device)
        arr = (C.c_void_p * world)(*self.peer_bases)
        h = C.c_void_p()
        check(self.l.idg_peers_create(self.base, self.nbytes, rank, world, arr, C.byref(h)), "idg_peers_create")
        self.handle = h
        if mc:
            check(self.l.idg_peers_set_multicast(h, mc), "idg_peers_set_multicast")
            self.multicast = True
        if os.environ.get("IDG_PEER_TIMEOUT_MS"):
            check(self.l.idg_peers_set_timeout_ms(h, int(os.environ["IDG_PEER_TIMEOUT_MS"])), "idg_peers_set_timeout_ms")
        self._off = 0
        self.state = self.carve((64,), torch.int32)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=group)

    def carve(self, shape, dtype=torch.float32):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        off = (self._off + 255) // 256 * 256
        if off + n > self.nbytes:
            raise RuntimeError("PeerSlab exhausted")
        self._off = off + n
        return self._bytes[off:off + n].view(dtype).view(*shape)

    def barrier(self):
        check(self.l.idg_peers_barrier(self.handle, ptr(self.state), cur_stream()), "idg_peers_barrier")

    def push(self, t):
        check(self.l.idg_peers_push(self.handle, ptr(t), t.numel() * t.element_size(), cur_stream()), "idg_peers_push")

    def status(self):
        """Synchronises the stream; raises IdgError (rc = IDG_ERR_PEER_TIMEOUT + rank) if a peer missed a barrier."""
        check(self.l.idg_peers_status(self.handle, ptr(self.state), cur_stream()), "idg_peers_status")


class DistFusedTrainer(EpochFetch):
    """LightGCN / SimGCL / XSimGCL training step, row-partitioned over `world` GPUs (same public surface as FusedTrainer).

    The contrastive models run their extra propagations the same way (local rows, finished rows stored to all peers by the
    SpMM epilogue, flag barrier per layer; sign-noise applied in the same epilogue from a noise tensor every rank draws
    identically), evaluate BPR + InfoNCE redundantly on the complete batch rows, add all row gradients into ONE table
    (the three propagations of SimGCL share their linear backward operator) and run one backward Horner chain whose
    last product applies Adam.  XSimGCL's captured-layer gradient joins the chain at that layer (idg_spmm_layer_add2)."""

    KINDS = ("LightGCN", "SimGCL", "XSimGCL")

    def __init__(self, kind, csr, table, num_users, K, reg_lambda, lr, rank, world, group=None, max_batch=1024,
                 betas=(0.9, 0.999), adam_eps=1e-8, use_cuda_graph=True, full_graph=None, closure_restrict="auto",
                 ssl_lambda=0.0, temperature=0.2, eps=0.0, cl_layer=1):
        if kind not in self.KINDS:
            raise NotImplementedError("row-partitioned training covers %s (got %s)" % (", ".join(self.KINDS), kind))
        if not 2 <= K <= 3:
            raise NotImplementedError("the row-partitioned step supports GCN_layer = 2 or 3 (got %d)" % K)
        if kind == "XSimGCL" and not 1 <= cl_layer <= K:
            raise NotImplementedError("row-partitioned XSimGCL needs 1 <= cl_layer <= GCN_layer")
        self.l = _lib.lib()
        self.kind, self.rank, self.world = kind, rank, world
        self.U, (self.N, self.d), self.K = num_users, table.shape, K
        self.reg_lambda, self.lr, self.betas, self.adam_eps = reg_lambda, lr, betas, adam_eps
        self.ssl_lambda, self.temperature, self.eps, self.cl_layer = ssl_lambda, temperature, eps, cl_layer
        self.inc0 = kind == "LightGCN"                     # layer 0 in the mean (LightGCN.py:41 vs SimGCL.py:45)
        self.cnt = float(K + (1 if self.inc0 else 0))
        dev = table.device
        self.dev = dev
        N, d = self.N, self.d
        nd = N * d * 4
        n_views = {"LightGCN": 1, "SimGCL": 3, "XSimGCL": 1}[kind]
        self.slab = PeerSlab((1 + (n_views + 1) * max(K - 1, 1)) * (nd + 4096) + (1 << 20), rank, world, group, dev)
        self.E0 = self.slab.carve((N, d))
        self.E0.copy_(table)
        # forward layer outputs X1..X_{K-1}, one set per propagation of the step (a peer may already be writing the next
        # propagation's layers while this rank still reads the previous one's in its batch-row layer)
        self.Wv = [[self.slab.carve((N, d)) for _ in range(K - 1)] for _ in range(n_views)]
        self.W = self.Wv[0]
        self.H = [self.slab.carve((N, d)) for _ in range(K - 1)]     # backward chain H_{K-1}..H_1
        self.bounds = partition_rows(csr.indptr.cpu().numpy(), world)
        self.b0, self.b1 = self.bounds[rank], self.bounds[rank + 1]
        self.local = Graph(csr, self.b0, self.b1)
        check(self.l.idg_graph_set_peers(self.local._h, self.slab.handle), "idg_graph_set_peers")
        self.full = full_graph if full_graph is not None else Graph(csr)   # row-restricted last layer + evaluation
        # Exchange of the FULL layers (forward layers, middle backward products, the Adam layer).  "fused": the propagation
        # kernel's epilogue stores every finished row to all peers.  "chunked": the local rows are computed in a few nnz-balanced
        # blocks with plain local stores and each finished block is streamed to the peers by a handful of CTAs on a second
        # stream while the next block is computed.  At the XL shape on 8 GPUs the fused form runs a layer over 1/8 of the rows in
        # 0.93 ms against 0.53 ms for the same layer without peer stores.  Measured (profiles/xl_exchange_r2.md): the chunked
        # form is 2 % faster there (3.37 vs 3.45 ms per step) -- the layer is bound by the multicast ingress itself (448 MB per
        # layer and GPU at ~480 GB/s), not by store back-pressure on the SMs; unicast peer stores are 15 % slower than multicast.
        # "auto" picks it only where it was measured to help: 8 ranks and a large local share (>= 4 M non-zeros per rank).
        mode = os.environ.get("IDG_DIST_EXCHANGE", "auto")
        local_nnz = int(csr.indptr[self.b1].item() - csr.indptr[self.b0].item())
        # (forced "chunked" also applies to a world of one -- the pushes are no-ops there -- so that the block schedule and its
        # fork / join are exercised on a single-GPU box)
        self.chunked = mode == "chunked" or (mode == "auto" and world >= 8 and local_nnz >= 4000000)
        self.chunks, self._push_side = [], None
        if self.chunked:
            n_chunks = max(1, int(os.environ.get("IDG_DIST_CHUNKS", "4")))
            self.push_ctas = max(1, int(os.environ.get("IDG_PUSH_CTAS", "16")))
            ip = csr.indptr[self.b0:self.b1 + 1].cpu().numpy().astype(np.int64)
            cuts = [self.b0]
            for c in range(1, n_chunks):
                r = self.b0 + int(np.searchsorted(ip, ip[0] + (ip[-1] - ip[0]) * c // n_chunks))
                cuts.append(min(max((r + 64) // 128 * 128, cuts[-1]), self.b1))
            cuts.append(self.b1)
            self.chunks = [(r0, r1, Graph(csr, r0, r1)) for r0, r1 in zip(cuts[:-1], cuts[1:]) if r1 > r0]
            self._push_side = torch.cuda.Stream(device=dev, priority=-1)
        z = lambda: torch.zeros(N, d, dtype=torch.float32, device=dev)
        self.gE0, self.m, self.v, self.G, self.F = z(), z(), z(), z(), z()
        self.rows = BatchRows(N, max_batch, dev)
        self.rows.worklist(self.full)
        contrastive = kind != "LightGCN"
        # batch-neighbourhood restriction of forward layer K-1 and of the first two backward products (see engine.py);
        # every rank derives the closure bitmap locally from the batch rows' neighbour lists (it holds the whole graph handle)
        if closure_restrict == "auto":
            from .graph import expected_closure_fraction
            closure_restrict = K >= 3 and not contrastive and expected_closure_fraction(csr, num_users, max_batch) < 0.4
        self.use_closure = bool(closure_restrict) and K >= 3 and not contrastive
        self.closure = self.slab.carve(((N + 127) // 128 * 4 + 4,), torch.int32)
        self.closure.zero_()
        if self.use_closure:
            self.rows.closure = self.closure
        self.max_batch, self.step_count = max_batch, 0
        self.ws = torch.empty(int(self.l.idg_bpr_workspace_bytes(max_batch)), dtype=torch.uint8, device=dev)
        self.n_loss = 3 if contrastive else 2
        self.loss = torch.zeros(4, dtype=torch.float32, device=dev)
        self.loss_acc = torch.zeros(4, dtype=torch.float64, device=dev)
        self.batch = torch.zeros(3, max_batch, dtype=torch.int64, device=dev)
        self.d_step = torch.zeros(1, dtype=torch.int32, device=dev)
        self._ep, self._ep_ptrs = None, torch.zeros(4, dtype=torch.int64, device=dev)     # registered epoch arrays (EpochFetch)
        self.regc = torch.zeros(N, dtype=torch.float32, device=dev)
        self.adam_scalars = torch.zeros(2, dtype=torch.float32, device=dev)
        self.adam_args = _lib.AdamArgs(ptr(self.E0), ptr(self.m), ptr(self.v), ptr(self.regc), ptr(self.adam_scalars), betas[0], betas[1], adam_eps)
        if contrastive:
            self.noise = torch.empty(K, N, d, dtype=torch.float32, device=dev)
            self.nce_ws = torch.empty(int(self.l.idg_infonce_workspace_bytes(max_batch, d)), dtype=torch.uint8, device=dev)
            self.nce_ws2 = torch.empty_like(self.nce_ws)            # item-side term on a second stream (engine.py:_contrast_pair)
            self._nce_side = torch.cuda.Stream(device=dev)
            self.V1 = z()
            self.V2 = z() if kind == "SimGCL" else None
            self.Gcl = z() if kind == "XSimGCL" else None
            self.Hk = z() if (kind == "XSimGCL" and cl_layer == K) else None
            self.uidx = torch.zeros(max_batch, dtype=torch.int64, device=dev)
            self.iidx = torch.zeros(max_batch, dtype=torch.int64, device=dev)
            self.ucnt = torch.zeros(1, dtype=torch.int32, device=dev)
            self.icnt = torch.zeros(1, dtype=torch.int32, device=dev)
        self.injected_noise = None   # parity tests: list of per-view [K,N,d] tensors
        self.use_cuda_graph = use_cuda_graph
        self._graphs, self._graph_launches, self.replayed_launches = {}, {}, 0
        self._prof = None
        torch.cuda.synchronize()
        self._host_barrier(group)
        self.group = group

    @staticmethod
    def _host_barrier(group):
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.barrier(group=group)

    def _exchanged(self, launch, out):
        """One full exchanged layer: ``launch(graph)`` enqueues the product over that handle's rows, writing rows of ``out`` (a
        slab tensor).  Fused mode: one launch over the local rows, peer stores in the epilogue.  Chunked mode: block by block,
        each finished block pushed from the side stream while the next one is computed; joined before returning."""
        if not self.chunked:
            launch(self.local)
            return
        main, side, d = torch.cuda.current_stream(), self._push_side, self.d
        for r0, r1, g in self.chunks:
            launch(g)
            side.wait_stream(main)
            check(self.l.idg_peers_push_ctas(self.slab.handle, out.data_ptr() + r0 * d * 4, (r1 - r0) * d * 4, self.push_ctas, side.cuda_stream),
                  "idg_peers_push_ctas")
        main.wait_stream(side)

    # ------------------------------------------------------------------
    def _mark(self, name):
        if self._prof is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._prof.append((name, e))

    def _draw_noise(self, view):
        """Every rank draws the SAME full noise tensor (same device-generator state on all ranks after tools.set_seed), in the
        reference's call pattern: one rand_like([N,d]) per layer, views in order (SimGCL.py:50)."""
        if self.injected_noise is not None:
            self.noise.copy_(self.injected_noise[view])
        else:
            for k in range(self.K):
                self.noise[k].uniform_()

    def _forward(self, W, out_mean, noise=False, out_cl=None):
        """One propagation: layers 1..K-1 on the local rows (rows pushed to every peer by the epilogue), last layer + layer
        mean on the batch rows only, every rank computing all of them (no exchange)."""
        l, s, K, d, loc, rows, slab = self.l, cur_stream(), self.K, self.d, self.local, self.rows, self.slab
        eps = self.eps if noise else 0.0
        x = self.E0
        for k in range(K - 1):
            nz = self.noise[k] if noise else None
            if self.use_closure and k == K - 2:
                # layer K-1 only on the batch rows and their neighbours: the restricted last layer reads nothing else
                check(l.idg_spmm_layer_masked(loc._h, ptr(x), ptr(W[k]), None, 0.0, None, None, 1.0, d, ptr(self.closure), s), "idg_spmm_layer_masked")
            else:
                self._exchanged(lambda g, x=x, k=k, nz=nz: g.spmm_layer(x, Y=W[k], noise=nz, eps=eps), W[k])
            self._mark('fwd_layer%d' % (k + 1))
            slab.barrier()
            self._mark('barrier')
            x = W[k]
        acc = ([self.E0] if self.inc0 else []) + list(W)
        acc += [None] * (3 - len(acc))
        y_cl = out_cl if (out_cl is not None and self.cl_layer == K) else None
        check(l.idg_spmm_layer_rows(self.full._h, ptr(x), ptr(y_cl), ptr(self.noise[K - 1]) if noise else None, eps, ptr(acc[0]), ptr(acc[1]), ptr(acc[2]),
                                    ptr(out_mean), self.cnt, d, ptr(rows.rowlist), ptr(rows.count), rows.max_rows, ptr(rows.worklist(self.full)), s),
              "idg_spmm_layer_rows")
        self._mark('fwd_last_rows')

    def _body(self, B, u, p, n):
        l, s, K, d = self.l, cur_stream(), self.K, self.d
        self._mark('start')
        loc, rows, slab = self.local, self.rows, self.slab
        contrastive = self.kind != "LightGCN"
        if contrastive:
            rows.build_unique(u, p, n, B, self.U, self.uidx, self.ucnt, self.iidx, self.icnt)
        else:
            rows.build(u, p, n, B, self.U)
        # the first backward product only publishes non-zero rows: this rank's copy of its output must be clear before the
        # peers write into it (they pass two barriers after this point in stream order first).  With the closure bitmap
        # the rows a step can write are known (batch rows + neighbours): they are re-zeroed at the END of the step
        # instead of the whole [N,d] buffer here (512 MB at the 1M x 1M size).
        if not self.use_closure:
            self.H[0].zero_()
        if self.use_closure:
            # every rank holds the whole graph handle (restricted last layer, evaluation): the closure of the batch is computed
            # locally from the batch rows' own neighbour lists -- identical on all ranks, no exchange, no barrier
            rows.build_closure(self.full)
        self._mark('batch_rows')
        cl_view = None
        if self.kind == "LightGCN":
            self._forward(self.Wv[0], self.F)
        elif self.kind == "SimGCL":
            self._forward(self.Wv[0], self.F)
            self._draw_noise(0)
            self._forward(self.Wv[1], self.V1, noise=True)
            self._draw_noise(1)
            self._forward(self.Wv[2], self.V2, noise=True)
        else:  # XSimGCL: one perturbed propagation, contrast view = post-noise output of layer cl_layer
            self._draw_noise(0)
            self._forward(self.Wv[0], self.F, noise=True, out_cl=self.V1)
            cl_view = self.V1 if self.cl_layer == K else self.Wv[0][self.cl_layer - 1]
        check(l.idg_bpr_forward(ptr(self.F), ptr(self.E0), u, p, n, B, self.U, self.N, d, self.reg_lambda, 7, ptr(self.loss), ptr(self.ws), s), "idg_bpr_forward")
        check(l.idg_bpr_backward(ptr(self.F), B, d, 7, None, ptr(self.G), self.reg_lambda, ptr(self.regc), ptr(self.ws), s), "idg_bpr_backward")
        check(l.idg_adam_prepare(ptr(self.d_step), ptr(self.adam_scalars), self.lr, self.betas[0], self.betas[1], s), "idg_adam_prepare")
        Gcl = None
        if contrastive:
            self.loss[2:3].zero_()
            Va, Vb, gA = (self.V1, self.V2, self.G) if self.kind == "SimGCL" else (cl_view, self.F, self.Gcl)
            main = torch.cuda.current_stream()
            self._nce_side.wait_stream(main)
            for (idx, cnt), ws, st in (((self.iidx, self.icnt), self.nce_ws2, self._nce_side), ((self.uidx, self.ucnt), self.nce_ws, main)):
                with torch.cuda.stream(st):     # disjoint gradient rows, separate workspaces: the two terms overlap
                    check(l.idg_infonce_fwd_bwd_dev(ptr(Va), ptr(Vb), ptr(idx), ptr(cnt), B, d, self.temperature, self.ssl_lambda,
                                                    ptr(self.loss[2:]), ptr(gA), ptr(self.G), ptr(ws), st.cuda_stream), "idg_infonce_fwd_bwd_dev")
            main.wait_stream(self._nce_side)
            Gcl = self.Gcl
        self._mark('bpr')
        # backward Horner chain on the local rows: H_K = G (+cnt Gcl if cl = K); H_l = G + A H_{l+1} (+cnt Gcl if cl = l);
        # the first product only gathers batch columns; the last product applies Adam in its epilogue and stores the updated
        # parameter rows to every peer: the gradient never goes to memory.
        import ctypes as _C
        adam = _C.byref(self.adam_args)
        h = self.G
        if Gcl is not None and self.cl_layer == K:
            check(l.idg_axpby(ptr(self.Hk), 1.0, ptr(self.G), self.cnt, ptr(Gcl), self.N * d, s), "idg_axpby")
            h = self.Hk
        for step_i in range(1, K):
            layer = K - step_i                                  # index of the H being produced
            out = self.H[step_i - 1]
            add2 = Gcl if (Gcl is not None and self.cl_layer == layer) else None
            if step_i == 1:
                if add2 is None and self.use_closure:   # only the batch neighbourhood can come out non-zero
                    check(l.idg_spmm_layer_sparse_in_masked(loc._h, ptr(h), ptr(out), ptr(self.G), d, ptr(rows.bitmap), ptr(self.closure), 1, s),
                          "idg_spmm_layer_sparse_in_masked")
                elif add2 is None:
                    check(l.idg_spmm_layer_sparse_in(loc._h, ptr(h), ptr(out), ptr(self.G), None, None, 1.0, d, ptr(rows.bitmap), 1, s), "idg_spmm_layer_sparse_in")
                else:   # every row is stored (no zero-row skipping with a second addend)
                    check(l.idg_spmm_layer_add2(loc._h, ptr(h), ptr(out), ptr(self.G), ptr(add2), self.cnt, d, ptr(rows.bitmap), 0, s), "idg_spmm_layer_add2")
                self._mark('bwd_sparse')
            elif self.use_closure and step_i == 2:  # H_{K-1} is zero outside the closure: gather only those columns
                check(l.idg_spmm_layer_sparse_in(loc._h, ptr(h), ptr(out), ptr(self.G), None, None, 1.0, d, ptr(self.closure), 0, s), "idg_spmm_layer_sparse_in")
                self._mark('bwd_layer')
            else:
                if add2 is None:
                    self._exchanged(lambda g, h=h, out=out: g.spmm_layer(h, Y=out, addend=self.G), out)
                else:
                    self._exchanged(lambda g, h=h, out=out, add2=add2: check(l.idg_spmm_layer_add2(g._h, ptr(h), ptr(out), ptr(self.G), ptr(add2), self.cnt, d,
                                                                                                 None, 0, cur_stream()), "idg_spmm_layer_add2"), out)
                self._mark('bwd_layer')
            slab.barrier()
            self._mark('barrier')
            h = out
        self._exchanged(lambda g, h=h: check(l.idg_spmm_layer_adam(g._h, ptr(h), ptr(self.G) if self.inc0 else None, self.cnt, d, adam, cur_stream()),
                                             "idg_spmm_layer_adam"), self.E0)
        self._mark('bwd_last_adam_push')
        check(l.idg_bpr_finish(ptr(self.E0), None, ptr(self.G), B, d, self.reg_lambda, None, ptr(self.regc), ptr(self.ws), s), "idg_bpr_finish")
        if Gcl is not None:
            for idx in (self.uidx, self.iidx):   # entries past the count are stale but valid rows of an all-zero table: harmless
                check(l.idg_zero_rows(ptr(Gcl), ptr(idx), B, d, s), "idg_zero_rows")
        if self.use_closure:
            check(l.idg_zero_rows_bitmap(ptr(self.H[0]), ptr(self.closure), self.N, d, s), "idg_zero_rows_bitmap")
        rows.clear()   # also zeroes the closure words (before the final barrier: peers publish theirs after it)
        self._mark('finish')
        slab.barrier()
        self._mark('barrier')
        check(l.idg_accumulate_f64(ptr(self.loss_acc), ptr(self.loss), 4, s), "idg_accumulate_f64")

    def step(self, users, pos, neg, apply_adam=True):
        assert apply_adam, "the distributed step always applies Adam"
        B = int(users.numel())
        # consecutive slices of the epoch's sample arrays are fetched by the captured step itself (engine.py:EpochFetch)
        fetch = self.use_cuda_graph and self._epoch_cursor(B, users, pos, neg)
        key = ("f", B, self._ep[2]) if fetch else B
        if not fetch:
            self.batch[0, :B].copy_(users); self.batch[1, :B].copy_(pos); self.batch[2, :B].copy_(neg)
        u, p, n = (self.batch[k].data_ptr() for k in range(3))
        if not self.use_cuda_graph:
            self._body(B, u, p, n)
        else:
            if key not in self._graphs:
                torch.cuda.synchronize()
                n0 = self.l.idg_launch_count()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    if fetch:
                        self._fetch_batch(B, key[2])
                    self._body(B, u, p, n)
                self._graph_launches[key] = int(self.l.idg_launch_count() - n0)
                self._graphs[key] = gr
            self._graphs[key].replay()
            self.replayed_launches += self._graph_launches[key]
        self.step_count += 1
        return self.loss[:self.n_loss]

    def pop_epoch_losses(self):
        self.slab.status()   # a peer that missed a flag barrier invalidates the epoch: surface it here, not as a hang
        out = self.loss_acc[:self.n_loss].tolist()
        self.loss_acc.zero_()
        return out

    @torch.no_grad()
    def final_embeddings(self):
        """Clean propagation of the current table on the whole graph (every rank, for its evaluation shard)."""
        return self.full.propagate_fwd(self.E0, self.K, self.inc0)

    def profile_steps(self, batches):
        """Eager steps with CUDA events between phases -> {phase: mean ms} (diagnostics)."""
        graph, self.use_cuda_graph = self.use_cuda_graph, False
        acc, n = {}, 0
        for b in batches:
            self._prof = []
            self.step(*b)
            torch.cuda.synchronize()
            for (n0, e0), (n1, e1) in zip(self._prof[:-1], self._prof[1:]):
                acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
            n += 1
        self._prof = None
        self.use_cuda_graph = graph
        return {k: v / n for k, v in acc.items()}
