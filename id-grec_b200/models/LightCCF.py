"""LightCCF (Zhang et al., SIGIR'25) -- same class interface as the reference's models/LightCCF.py:16-120:
LightGCN (or MF) encoder, BPR + ego L2 + the neighbourhood-aggregation contrastive loss over the batch."""
import utility.utility_data.data_graph
import utility.utility_train.trainer as trainer
from idgrec import ops
from idgrec.model_base import PropagationModel


class LightCCF(PropagationModel):
    kind = "LightCCF"
    graph_capturable = True   # forward() has no host sync: universal_trainer replays the whole step from a CUDA graph

    def __init__(self, config, dataset, device):
        super(LightCCF, self).__init__(config, dataset, device, utility.utility_data.data_graph.sparse_adjacency_matrix)
        self.ssl_lambda = float(config['ssl_lambda'])
        self.temperature = float(config['temperature'])

    def fused_trainer(self, lr, max_batch):
        """LightGCN encoder: the fused CUDA-graph step of idgrec.engine (row-restricted propagation, batch x batch loss on
        tensor cores, Adam in the last backward epilogue).  MF encoder: autograd ops + torch.optim.Adam."""
        if self.config['encoder'] == 'MF':
            return None
        return super(LightCCF, self).fused_trainer(lr, max_batch)

    def aggregate(self):
        """LightCCF.py:44-62 (the MF encoder returns the ego tables, :59-60)."""
        return self._split(self.encode())

    def get_neighbor_aggregate_loss(self, embedding1, embedding2, tau):
        """LightCCF.py:81-94 on dense [B,d] blocks."""
        return ops.pair_loss("lightccf", embedding1, embedding2, tau)

    def forward(self, user, positive, negative):
        """LightCCF.py:58-79 -> [bpr, reg_lambda * reg, ssl_lambda * na]."""
        E0 = self.table()
        final = self.encode(E0)
        loss = ops.bpr_reg_loss(final, E0, user, positive, negative, self.dataset.num_users, self.reg_lambda, 7)
        ue, pe = self.batch_rows(final, user, positive)
        na_loss = self.ssl_lambda * self.get_neighbor_aggregate_loss(ue, pe, self.temperature)
        return [loss[0], loss[1], na_loss]


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = LightCCF(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger

    def train(self):
        trainer.universal_trainer(self.model, self.args, self.config, self.dataset, self.device, self.logger)
