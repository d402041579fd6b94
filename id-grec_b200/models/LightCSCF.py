"""LightCSCF -- same class interface as the reference's models/LightCSCF.py:14-129: LightGCN (or MF) encoder,
ego L2 + the margin variant of the neighbourhood-aggregation loss (plus BPR with the MF encoder)."""
import utility.utility_data.data_graph
import utility.utility_train.trainer as trainer
from idgrec import ops
from idgrec.model_base import PropagationModel


class LightCSCF(PropagationModel):
    kind = "LightCSCF"
    graph_capturable = True   # forward() has no host sync: universal_trainer replays the whole step from a CUDA graph

    def __init__(self, config, dataset, device):
        super(LightCSCF, self).__init__(config, dataset, device, utility.utility_data.data_graph.sparse_adjacency_matrix)
        self.temperature = float(config['temperature'])
        self.lambda_gamma = float(config['lambda_gamma'])
        self.lambda_reg = float(config['lambda_reg'])
        self.lambda_margin = float(config['lambda_margin'])

    def fused_trainer(self, lr, max_batch):
        """LightGCN encoder: the fused CUDA-graph step of idgrec.engine (row-restricted propagation, batch x batch loss on
        tensor cores, Adam in the last backward epilogue).  MF encoder: autograd ops + torch.optim.Adam."""
        if self.config['encoder'] == 'MF':
            return None
        return super(LightCSCF, self).fused_trainer(lr, max_batch)

    def aggregate(self):
        return self._split(self.encode())

    def LightCSCF_loss(self, embedding1, embedding2, temperature):
        """LightCSCF.py:93-104 on dense [B,d] blocks."""
        return ops.pair_loss("lightcscf", embedding1, embedding2, temperature, self.lambda_margin)

    def forward(self, user, positive, negative):
        """LightCSCF.py:58-91 -> [bpr, reg, na] with the MF encoder, [reg, na] with LightGCN."""
        E0 = self.table()
        final = self.encode(E0)
        loss = ops.bpr_reg_loss(final, E0, user, positive, negative, self.dataset.num_users, self.lambda_reg, 7)
        ue, pe = self.batch_rows(final, user, positive)
        na_loss = self.lambda_gamma * self.LightCSCF_loss(ue, pe, self.temperature)
        if self.config['encoder'] == 'MF':
            return [loss[0], loss[1], na_loss]
        return [loss[1], na_loss]


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = LightCSCF(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger

    def train(self):
        trainer.universal_trainer(self.model, self.args, self.config, self.dataset, self.device, self.logger)
