"""SimGCL (Yu et al., SIGIR'22) -- same class interface as the reference's models/SimGCL.py:14-113:
layer mean without layer 0, sign-noise perturbed views, in-batch InfoNCE on the batch's unique
users / positive items."""
import torch

import utility.utility_data.data_graph
import utility.utility_train.trainer as trainer
from idgrec import ops
from idgrec.model_base import PropagationModel


class SimGCL(PropagationModel):
    kind = "SimGCL"

    def __init__(self, config, dataset, device):
        super(SimGCL, self).__init__(config, dataset, device, utility.utility_data.data_graph.sparse_adjacency_matrix)
        self.ssl_lambda = float(config['ssl_lambda'])
        self.epsilon = float(config['epsilon'])
        self.temperature = float(config['temperature'])

    def _noise(self, E0):
        # SimGCL.py:50: one torch.rand_like([N,d]) per layer from the device generator
        return torch.stack([torch.rand_like(E0) for _ in range(self.num_layers)])

    def _propagate(self, E0, perturbed, noise=None):
        if perturbed and noise is None:
            noise = self._noise(E0.detach())
        return ops.propagate(E0, self.Graph, self.num_layers, include_layer0=False,
                             noise=noise if perturbed else None, eps=self.epsilon if perturbed else 0.0)

    def aggregate(self, perturbed=False, noise=None):
        """SimGCL.py:39-60."""
        return self._split(self._propagate(self.table(), perturbed, noise))

    def forward(self, user, positive, negative, noises=None):
        """SimGCL.py:62-90 -> [bpr, reg, ssl].  ``noises`` = (view1, view2) stacks of K [N,d] tensors
        lets parity tests inject the reference's draws."""
        U = self.dataset.num_users
        E0 = self.table()
        final = self._propagate(E0, False)
        v1 = self._propagate(E0, True, None if noises is None else noises[0])
        v2 = self._propagate(E0, True, None if noises is None else noises[1])
        loss = ops.bpr_reg_loss(final, E0, user, positive, negative, U, self.reg_lambda, 7)
        user_index = torch.unique(user)
        item_index = torch.unique(positive) + U
        ssl = ops.infonce_rows(v1, v2, user_index, self.temperature) + ops.infonce_rows(v1, v2, item_index, self.temperature)
        return [loss[0], loss[1], self.ssl_lambda * ssl]


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = SimGCL(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger

    def train(self):
        trainer.universal_trainer(self.model, self.args, self.config, self.dataset, self.device, self.logger)
