"""utility_data/data_graph.py of the reference (data_graph.py:7-55): the symmetric-normalised
adjacency, built on the device as canonical CSR by libidgrec_sm100.so instead of scipy dok/lil
slicing (89 s on the yelp2018 shape).  Values are bit-identical to the reference's; no ``pre_A.npz``
cache is written or trusted (the reference loads a stale cache blindly)."""
import torch

from idgrec.graph import Graph, build_norm_adjacency


class NormAdjacency:
    """What sparse_adjacency_matrix returns.  Plays the role of the scipy matrix *and* of the torch
    COO tensor of the reference: ``tools.convert_sp_mat_to_sp_tensor(adj).coalesce().to(device)``
    (models/LightGCN.py:30-32) yields the propagation handle used in place of torch.sparse.mm."""

    def __init__(self, data, add_self):
        self.data, self.add_self = data, add_self
        self._graphs = {}
        self.shape = (data.num_nodes, data.num_nodes)

    def coalesce(self):
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ID-GRec B200 hot path needs a CUDA device (got %s); there is no CPU fallback" % device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = str(device)
        if key not in self._graphs:
            d = self.data
            csr = build_norm_adjacency(d.train_user, d.train_item, d.num_users, d.num_items, add_self=self.add_self, device=device)
            self._graphs[key] = Graph(csr)
        return self._graphs[key]

    def tocsr(self, device="cuda"):
        return self.to(device).csr.to_scipy()

    def tocoo(self, device="cuda"):
        return self.tocsr(device).tocoo()


class EdgeListAdjacency(NormAdjacency):
    """D^-1/2 A D^-1/2 of an explicit (user, item) edge list -- the edge-dropped sub-graphs of SGL
    (utility_function/tools.py:67-92) go through the same device CSR builder as the full graph."""

    def __init__(self, user_index, item_index, num_users, num_items):
        class _Edges:
            pass
        d = _Edges()
        d.train_user, d.train_item = user_index, item_index
        d.num_users, d.num_items, d.num_nodes = int(num_users), int(num_items), int(num_users) + int(num_items)
        super().__init__(d, add_self=False)


class BipartiteAdjacency(NormAdjacency):
    """What sparse_adjacency_matrix_R returns: D_u^-1/2 R D_i^-1/2 (float64 arithmetic, rounded to fp32 once by
    tools.py:101).  On the device it is kept as the symmetric [[0, R_hat], [R_hat^T, 0]] CSR; ``.to(device)`` yields a
    handle with the user-row half (R_hat . x) and the item-row half (R_hat^T . x) as two propagation graphs."""

    class Handle:
        def __init__(self, csr, num_users):
            self.csr = csr
            self.users = Graph(csr, 0, num_users)                 # rows of R_hat:   user_emb = R_hat . item_emb
            self.items = Graph(csr, num_users, csr.shape[0])      # rows of R_hat^T: item_emb = R_hat^T . user_emb
            self.num_users = num_users

    def __init__(self, data):
        super().__init__(data, add_self=False)

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ID-GRec B200 hot path needs a CUDA device (got %s); there is no CPU fallback" % device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = str(device)
        if key not in self._graphs:
            d = self.data
            csr = build_norm_adjacency(d.train_user, d.train_item, d.num_users, d.num_items, add_self=False, device=device, f64_degrees=True)
            self._graphs[key] = BipartiteAdjacency.Handle(csr, d.num_users)
        return self._graphs[key]

    def tocsr(self, device="cuda"):
        """The [U, I] matrix itself (scipy), e.g. for comparison with the reference's pre_R.npz."""
        h = self.to(device)
        full = h.csr.to_scipy()
        return full[:h.num_users, h.num_users:].tocsr()


def sparse_adjacency_matrix_R(data):
    """D_u^-1/2 R D_i^-1/2 (data_graph.py:56-77; EGCF).  No pre_R.npz cache is written or trusted."""
    return BipartiteAdjacency(data)


def sparse_adjacency_matrix_with_self(data):
    """D^-1/2 (A + I) D^-1/2, float64 arithmetic rounded to fp32 once (data_graph.py:7-30; NGCF)."""
    return NormAdjacency(data, add_self=True)


def sparse_adjacency_matrix(data):
    """D^-1/2 A D^-1/2 in float32 (data_graph.py:33-55; LightGCN / SimGCL / XSimGCL)."""
    return NormAdjacency(data, add_self=False)
