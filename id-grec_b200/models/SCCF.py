"""SCCF (Wu et al., KDD'24) -- same class interface as the reference's models/SCCF.py:15-104: LightGCN (or MF)
encoder and the two-term contrastive objective [-up, down]."""
import torch

import utility.utility_data.data_graph
import utility.utility_train.trainer as trainer
from idgrec import ops
from idgrec.model_base import PropagationModel


class SCCF(PropagationModel):
    kind = "SCCF"

    def __init__(self, config, dataset, device):
        super(SCCF, self).__init__(config, dataset, device, utility.utility_data.data_graph.sparse_adjacency_matrix)
        self.temperature = float(config['temperature'])
        self.encoder = config['encoder']

    def fused_trainer(self, lr, max_batch):
        """LightGCN encoder: the fused CUDA-graph step of idgrec.engine (row-restricted propagation, batch x batch loss on
        tensor cores, Adam in the last backward epilogue).  MF encoder: autograd ops + torch.optim.Adam."""
        if self.config['encoder'] == 'MF':
            return None
        return super(SCCF, self).fused_trainer(lr, max_batch)

    def aggregate(self):
        return self._split(self.encode())

    def forward(self, user, positive, negative):
        """SCCF.py:54-81.  ``down`` sums psi over ALL batch pairs: each (unique user, unique item) pair then appears
        count_u * count_i times, which is the reference's count-weighted sum over the unique ids (SCCF.py:72-79);
        only the two unique counts are needed for the mean."""
        final = self.encode()
        ue, pe = self.batch_rows(final, user, positive)
        up_neg = ops.pair_loss("sccf_up", ue, pe, self.temperature)
        denom = float(torch.unique(user).numel()) * float(torch.unique(positive).numel())
        down = ops.pair_loss("sccf_down", ue, pe, self.temperature, denom)
        return [up_neg, down]


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = SCCF(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger

    def train(self):
        trainer.universal_trainer(self.model, self.args, self.config, self.dataset, self.device, self.logger)
