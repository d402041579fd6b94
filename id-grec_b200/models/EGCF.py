"""EGCF (Zhang et al., TOIS'24) -- same class interface as the reference's models/EGCF.py:14-138: an embedding-less
user side (users are tanh(R_hat . item_embedding)), tanh between the propagation layers, layer SUM instead of mean,
BPR + item ego L2 + three in-batch InfoNCE terms (user-user, item-item, user-item) on the batch rows."""
import torch
from torch import nn

import utility.utility_data.data_graph
import utility.utility_function.tools as tools
import utility.utility_train.trainer as trainer
from idgrec import ops


class EGCF(nn.Module):
    kind = "EGCF"
    fused_trainer = None       # autograd ops over the CUDA kernels + torch.optim.Adam (reference loop trainer.py:40-56)
    graph_capturable = True    # forward() has no host sync: universal_trainer replays the whole step from a CUDA graph

    def __init__(self, config, dataset, device):
        from idgrec.model_base import validate_config
        validate_config("EGCF", config, dataset)
        super(EGCF, self).__init__()
        self.config, self.dataset, self.device = config, dataset, device
        self.reg_lambda = float(config['reg_lambda'])
        self.ssl_lambda = float(config['ssl_lambda'])
        self.temperature = float(config['temperature'])
        self.aggregate_mode = config['mode']
        self.user_embedding = None
        # same torch-generator draws as EGCF.py:27-31: Embedding's own normal_ init, then xavier on the item table only
        self.item_embedding = nn.Embedding(num_embeddings=dataset.num_items, embedding_dim=int(config['embedding_size']))
        nn.init.xavier_uniform_(self.item_embedding.weight, gain=1)
        self.user_Graph = utility.utility_data.data_graph.sparse_adjacency_matrix_R(dataset)
        self.user_Graph = tools.convert_sp_mat_to_sp_tensor(self.user_Graph)
        self.user_Graph = self.user_Graph.coalesce().to(self.device)
        self.Graph = None
        if self.aggregate_mode == 'parallel':
            self.Graph = utility.utility_data.data_graph.sparse_adjacency_matrix(dataset)
            self.Graph = tools.convert_sp_mat_to_sp_tensor(self.Graph)
            self.Graph = self.Graph.coalesce().to(self.device)
        self.activation_layer = nn.Tanh()
        self.activation = nn.Sigmoid()

    # node tables are [N, d] with the user rows first (as everywhere on this path); a side that does not exist yet is zero
    def _pad_items(self, item_rows):
        U = self.dataset.num_users
        return torch.cat([torch.zeros((U, item_rows.shape[1]), dtype=item_rows.dtype, device=item_rows.device), item_rows])

    def _users_from_items(self, X):
        """tanh(R_hat . item rows of X) in the user rows, zeros in the item rows (EGCF.py:52 / :67)."""
        return ops.tanh(ops.spmm_rows(X, self.user_Graph.users, self.user_Graph.items))

    def _items_from_users(self, X):
        """tanh(R_hat^T . user rows of X) in the item rows (EGCF.py:53)."""
        return ops.tanh(ops.spmm_rows(X, self.user_Graph.items, self.user_Graph.users))

    def _final(self):
        U, K = self.dataset.num_users, int(self.config['GCN_layer'])
        items0 = self._pad_items(self.item_embedding.weight)
        if self.aggregate_mode == 'parallel':
            # EGCF.py:64-84: users from one R_hat product, then K tanh(A_hat .) layers on [users; items], summed
            x = self._users_from_items(items0) + items0
            total = None
            for _ in range(K):
                x = ops.tanh(ops.spmm(x, self.Graph))
                total = x if total is None else total + x
            return total
        # EGCF.py:45-62: alternate users <- items, items <- users; sum each side over the layers
        cur, total = items0, None
        for _ in range(K):
            users = self._users_from_items(cur)
            cur = self._items_from_users(users)
            layer = users + cur
            total = layer if total is None else total + layer
        return total

    def aggregate(self):
        final = self._final()
        return torch.split(final, [self.dataset.num_users, self.dataset.num_items])

    def parallel_aggregate(self):
        return self.aggregate()

    def alternating_aggregate(self):
        return self.aggregate()

    def forward(self, user, positive, negative):
        """EGCF.py:86-112 -> [bpr, reg_lambda * reg(ego_pos, ego_neg), ssl_lambda * (nce(u,u) + nce(p,p) + nce(u,p))]."""
        U = self.dataset.num_users
        final = self._final()
        ego = self._pad_items(self.item_embedding.weight)
        loss = ops.bpr_reg_loss(final, ego, user, positive, negative, U, self.reg_lambda, 6)
        ue = ops.gather_rows(final, user.long())
        pe = ops.gather_rows(final, positive.long() + U)
        rows = torch.arange(user.numel(), device=final.device)
        ssl = ops.infonce_rows(ue, ue, rows, self.temperature) + ops.infonce_rows(pe, pe, rows, self.temperature) \
            + ops.infonce_rows(ue, pe, rows, self.temperature)
        return [loss[0], loss[1], self.ssl_lambda * ssl]

    def final_embeddings(self):
        with torch.no_grad():
            return self.aggregate()

    def get_rating_for_test(self, user):
        with torch.no_grad():
            users_emb, items_emb = self.aggregate()
            return ops.rating_matrix(users_emb.contiguous(), items_emb.contiguous(), user)


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = EGCF(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger

    def train(self):
        trainer.universal_trainer(self.model, self.args, self.config, self.dataset, self.device, self.logger)
