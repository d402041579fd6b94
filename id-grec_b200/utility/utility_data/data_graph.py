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


def sparse_adjacency_matrix_with_self(data):
    """D^-1/2 (A + I) D^-1/2, float64 arithmetic rounded to fp32 once (data_graph.py:7-30; NGCF)."""
    return NormAdjacency(data, add_self=True)


def sparse_adjacency_matrix(data):
    """D^-1/2 A D^-1/2 in float32 (data_graph.py:33-55; LightGCN / SimGCL / XSimGCL)."""
    return NormAdjacency(data, add_self=False)
