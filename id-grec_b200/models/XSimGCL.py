"""XSimGCL (Yu et al., TKDE'23) -- same class interface as the reference's models/XSimGCL.py:14-118:
one perturbed propagation per step, the contrast view is the post-noise output of layer
``cl_layer``; evaluation is unperturbed."""
import torch

import utility.utility_data.data_graph
import utility.utility_train.trainer as trainer
from idgrec import ops
from idgrec.model_base import PropagationModel


class XSimGCL(PropagationModel):
    kind = "XSimGCL"

    def __init__(self, config, dataset, device):
        super(XSimGCL, self).__init__(config, dataset, device, utility.utility_data.data_graph.sparse_adjacency_matrix)
        self.ssl_lambda = float(config['ssl_lambda'])
        self.epsilon = float(config['epsilon'])
        self.temperature = float(config['temperature'])
        self.cl_layer = int(config['cl_layer'])

    def aggregate(self, perturbed=False, noise=None):
        """XSimGCL.py:40-67: perturbed -> (users, items, users_cl, items_cl); else (users, items)."""
        E0 = self.table()
        K = self.num_layers
        if not perturbed:
            return self._split(ops.propagate(E0, self.Graph, K, include_layer0=False))
        if noise is None:
            noise = torch.stack([torch.rand_like(E0.detach()) for _ in range(K)])
        if 1 <= self.cl_layer <= K:
            final, cl = ops.propagate(E0, self.Graph, K, include_layer0=False, noise=noise, eps=self.epsilon, cl_layer=self.cl_layer)
        else:  # XSimGCL.py:48: the contrast view stays the ego table
            final, cl = ops.propagate(E0, self.Graph, K, include_layer0=False, noise=noise, eps=self.epsilon), E0
        return (*self._split(final), *self._split(cl))

    def forward(self, user, positive, negative, noise=None):
        """XSimGCL.py:69-95 -> [bpr, reg, ssl]."""
        U = self.dataset.num_users
        E0 = self.table()
        K = self.num_layers
        if noise is None:
            noise = torch.stack([torch.rand_like(E0.detach()) for _ in range(K)])
        if 1 <= self.cl_layer <= K:
            final, cl = ops.propagate(E0, self.Graph, K, include_layer0=False, noise=noise, eps=self.epsilon, cl_layer=self.cl_layer)
        else:  # XSimGCL.py:48: no layer is captured, the contrast view stays the ego table
            final, cl = ops.propagate(E0, self.Graph, K, include_layer0=False, noise=noise, eps=self.epsilon), E0
        loss = ops.bpr_reg_loss(final, E0, user, positive, negative, U, self.reg_lambda, 7)
        user_index = torch.unique(user)
        item_index = torch.unique(positive) + U
        ssl = ops.infonce_rows(cl, final, user_index, self.temperature) + ops.infonce_rows(cl, final, item_index, self.temperature)
        return [loss[0], loss[1], self.ssl_lambda * ssl]


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = XSimGCL(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger

    def train(self):
        trainer.universal_trainer(self.model, self.args, self.config, self.dataset, self.device, self.logger)
