"""Shared machinery of the propagation models (the per-model files under models/ keep the
reference's class names, constructor and method signatures; SURVEY.md section 8 b)."""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .engine import FusedTrainer


def validate_config(kind, config, dataset):
    """Checks a configuration against the limits of the CUDA kernels BEFORE training starts (the reference accepts any
    embedding_size / GCN_layer / top_K; a value outside the kernels must not surface as an error inside the first
    Test() after a whole epoch).  Evaluation itself takes any embedding width and any top_K (idg_eval_topk ranks
    exhaustively outside its tiled shapes)."""
    problems = []
    d = int(config['embedding_size'])
    mf = kind == "MFBPR" or str(config.get('encoder', 'LightGCN')) == 'MF'
    if not mf and d not in (32, 64, 128):
        problems.append("embedding_size = %d: the propagation kernel handles 32, 64 or 128" % d)
    if mf and d not in (32, 64, 128, 192, 256, 320):
        problems.append("embedding_size = %d: the BPR kernels handle 32, 64, 128, 192, 256 or 320" % d)
    if kind in ("SimGCL", "XSimGCL", "SGL", "EGCF") and d != 64:
        problems.append("embedding_size = %d: the InfoNCE kernels of %s handle 64" % (d, kind))
    if kind == "NGCF":
        layers = eval(config['layer_size'])
        K = int(config['GCN_layer'])
        if d != 64 or any(int(x) != 64 for x in layers[:K]):
            problems.append("NGCF: embedding_size and every used layer_size must be 64 (dense-layer kernels are 64 x 64)")
        if K not in (1, 2, 3, 4):
            problems.append("NGCF: GCN_layer = %d (the concatenated width 64 (K+1) must be 128, 192, 256 or 320)" % K)
        if len(layers) < K or len(eval(config.get('mess_drop_prob', '[]'))) < (K if eval(config.get('mess_dropout', 'False')) else 0):
            problems.append("NGCF: layer_size / mess_drop_prob shorter than GCN_layer")
    elif not mf and int(config.get('GCN_layer', 1)) < 1:
        problems.append("GCN_layer must be >= 1")
    try:
        topk = [int(k) for k in eval(config['top_K'])]
    except Exception:  # noqa: BLE001
        topk = None
    if not topk or any(k < 1 for k in topk):
        problems.append("top_K = %r: a non-empty list of positive cut-offs is needed" % (config.get('top_K'),))
    elif len(topk) > 8:
        problems.append("top_K has %d cut-offs: the metric kernel takes at most 8 per evaluation" % len(topk))
    elif max(topk) > dataset.num_items:
        problems.append("top_K = %s exceeds the number of items (%d)" % (topk, dataset.num_items))
    elif max(topk) > 256:
        problems.append("top_K = %s: the ranking kernels keep at most 256 entries per user" % (topk,))
    if problems:
        raise ValueError("configuration outside the accelerated path of %s:\n  - %s" % (kind, "\n  - ".join(problems)))


class PropagationModel(nn.Module):
    """nn.Module with the reference's duck-typed model contract:
    ``forward(user, pos, neg) -> [losses]``, ``aggregate(...) -> (users_emb, items_emb)``,
    ``get_rating_for_test(user) -> [b, I]``; attributes Graph, user_embedding, item_embedding,
    activation, config, dataset, device."""

    kind = "LightGCN"
    include_layer0 = True

    def __init__(self, config, dataset, device, graph_fn=None):
        super().__init__()
        validate_config(self.kind, config, dataset)
        self.config, self.dataset, self.device = config, dataset, device
        self.reg_lambda = float(config.get('reg_lambda', config.get('lambda_reg', 0.0)))
        dim = int(config['embedding_size'])
        self.user_embedding = nn.Embedding(num_embeddings=dataset.num_users, embedding_dim=dim)
        self.item_embedding = nn.Embedding(num_embeddings=dataset.num_items, embedding_dim=dim)
        # same draw order from the torch CPU generator as the reference (LightGCN.py:27-28)
        nn.init.xavier_uniform_(self.user_embedding.weight, gain=1)
        nn.init.xavier_uniform_(self.item_embedding.weight, gain=1)
        self.activation = nn.Sigmoid()
        self._table = None
        self._fused = None
        self.Graph = None
        if graph_fn is not None:
            import utility.utility_function.tools as tools
            self.Graph = graph_fn(dataset)
            self.Graph = tools.convert_sp_mat_to_sp_tensor(self.Graph)
            self.Graph = self.Graph.coalesce().to(self.device)

    @property
    def num_layers(self):
        return int(self.config.get('GCN_layer', 0))

    # -- one [N,d] table behind both embedding weights (no torch.cat per step) -----------------
    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self._fuse_tables()
        return self

    def _fuse_tables(self):
        uw, iw = self.user_embedding.weight, self.item_embedding.weight
        if not uw.is_cuda:
            self._table = None
            return
        U = uw.shape[0]
        t = self._table
        if t is not None and uw.data_ptr() == t.data_ptr() and iw.data_ptr() == t[U:].data_ptr():
            return
        t = torch.cat([uw.data, iw.data]).contiguous()
        uw.data, iw.data = t[:U], t[U:]
        self._table, self._fused = t, None

    def table(self):
        """[N,d] ego embeddings: the fused buffer on CUDA, an autograd-tracked cat when grads are needed."""
        if torch.is_grad_enabled() and self.user_embedding.weight.requires_grad:
            return torch.cat([self.user_embedding.weight, self.item_embedding.weight])
        if self._table is None:
            raise RuntimeError("model parameters are not on a CUDA device; the ID-GRec hot path has no CPU fallback")
        return self._table

    def _split(self, X):
        return torch.split(X, [self.dataset.num_users, self.dataset.num_items])

    # -- fused trainer (used by utility_train.trainer.universal_trainer) -------------------------
    def fused_trainer(self, lr, max_batch):
        if self.kind == "XSimGCL" and not 1 <= int(self.config.get('cl_layer', 1)) <= self.num_layers:
            return None   # contrast view = ego table (XSimGCL.py:48): the autograd ops in forward() cover it
        if self._fused is None or self._fused.lr != lr or self._fused.max_batch < max_batch:
            if self._table is None:
                self._fuse_tables()
            if self._table is None:
                raise RuntimeError("model must be moved to a CUDA device first (model.to(device))")
            cfg = self.config
            import torch.distributed as dist
            world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
            if world > 1 and int(cfg.get('num_gpus', world)) > 1 and self.kind in ("LightGCN", "SimGCL", "XSimGCL"):
                # row-partitioned over the process group: the parameter table moves into the peer slab
                from .dist import DistFusedTrainer
                ft = DistFusedTrainer(self.kind, self.Graph.csr, self._table, self.dataset.num_users, self.num_layers, self.reg_lambda,
                                      lr, dist.get_rank(), world, max_batch=max_batch, full_graph=self.Graph,
                                      use_cuda_graph=str(cfg.get('cuda_graph', '1')) not in ('0', 'False', 'false'),
                                      closure_restrict={'0': False, '1': True}.get(str(cfg.get('closure_restrict', 'auto')), 'auto'),
                                      ssl_lambda=float(cfg.get('ssl_lambda', 0.0)), temperature=float(cfg.get('temperature', 0.2)),
                                      eps=float(cfg.get('epsilon', 0.0)), cl_layer=int(cfg.get('cl_layer', 1)))
                U = self.dataset.num_users
                self._table = ft.E0
                self.user_embedding.weight.data, self.item_embedding.weight.data = ft.E0[:U], ft.E0[U:]
                self._fused = ft
                return ft
            self._fused = FusedTrainer(
                self.kind, self.Graph, self._table, self.dataset.num_users, self.num_layers, self.reg_lambda, lr,
                ssl_lambda=float(cfg.get('ssl_lambda', cfg.get('lambda_gamma', 0.0))), temperature=float(cfg.get('temperature', 0.2)),
                margin=float(cfg.get('lambda_margin', 0.0)), gamma=float(cfg.get('gamma', 0.0)),
                eps=float(cfg.get('epsilon', 0.0)), cl_layer=int(cfg.get('cl_layer', 1)), max_batch=max_batch,
                use_cuda_graph=str(cfg.get('cuda_graph', '1')) not in ('0', 'False', 'false'),
                restrict_rows=str(cfg.get('restrict_rows', '1')) not in ('0', 'False', 'false'),
                fuse_adam=str(cfg.get('fuse_adam', '1')) not in ('0', 'False', 'false'),
                closure_restrict={'0': False, '1': True}.get(str(cfg.get('closure_restrict', 'auto')), 'auto'))
        return self._fused

    # -- LightGCN-or-MF encoder of the loss-only models (LightCCF.py:59-62, DirectAU.py:60-66, ...) -----------
    def encode(self, E0=None):
        """[N,d] final table: K-layer propagation with the layer-0 mean, or the ego table itself for ``encoder = MF``."""
        E0 = self.table() if E0 is None else E0
        if str(self.config.get('encoder', 'LightGCN')) == 'MF':
            return E0
        return ops.propagate(E0, self.Graph, self.num_layers, include_layer0=True)

    def batch_rows(self, final, user, positive):
        """(F_u[user], F_i[positive]) as dense [B,d] blocks whose backward scatters deterministically."""
        return ops.gather_rows(final, user.long()), ops.gather_rows(final, positive.long() + self.dataset.num_users)

    # -- evaluation ------------------------------------------------------------------------------
    def final_embeddings(self):
        """(users_emb, items_emb) for evaluation: clean propagation, no grad."""
        with torch.no_grad():
            return self.aggregate()[:2]

    def get_rating_for_test(self, user):
        """LightGCN.py:74-80: sigmoid(U_b . F_i^T) as a dense [b, I] matrix.  Kept for API
        completeness; Test() ranks through the fused top-K kernel and never materialises it."""
        with torch.no_grad():
            users_emb, items_emb = self.aggregate()[:2]
            return ops.rating_matrix(users_emb, items_emb, user)
