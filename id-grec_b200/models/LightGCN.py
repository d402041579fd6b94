"""LightGCN (He et al., SIGIR'20) on the B200 hot path -- same class interface as the reference's
models/LightGCN.py:14-95: K-layer propagation, mean over the K+1 layer outputs, BPR + ego L2."""
import torch

import utility.utility_data.data_graph
import utility.utility_train.trainer as trainer
from idgrec import ops
from idgrec.model_base import PropagationModel


class LightGCN(PropagationModel):
    kind = "LightGCN"

    def __init__(self, config, dataset, device):
        super(LightGCN, self).__init__(config, dataset, device, utility.utility_data.data_graph.sparse_adjacency_matrix)

    def aggregate(self):
        """LightGCN.py:36-52 -> (users_emb [U,d], items_emb [I,d]); one fused K-layer kernel chain."""
        final = ops.propagate(self.table(), self.Graph, self.num_layers, include_layer0=True)
        return self._split(final)

    def forward(self, user, positive, negative):
        """LightGCN.py:54-72 -> [bpr_loss, reg_lambda * reg_loss] (autograd-capable)."""
        E0 = self.table()
        final = ops.propagate(E0, self.Graph, self.num_layers, include_layer0=True)
        loss = ops.bpr_reg_loss(final, E0, user, positive, negative, self.dataset.num_users, self.reg_lambda, 7)
        return [loss[0], loss[1]]


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = LightGCN(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger

    def train(self):
        trainer.universal_trainer(self.model, self.args, self.config, self.dataset, self.device, self.logger)
