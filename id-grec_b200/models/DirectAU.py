"""DirectAU (Wang et al., KDD'22) -- same class interface as the reference's models/DirectAU.py:15-105: LightGCN (or
MF) encoder, alignment + uniformity of the normalised batch rows + ego L2 (negatives are not used)."""
import utility.utility_data.data_graph
import utility.utility_train.trainer as trainer
from idgrec import ops
from idgrec.model_base import PropagationModel


class DirectAU(PropagationModel):
    kind = "DirectAU"
    graph_capturable = True   # forward() has no host sync: universal_trainer replays the whole step from a CUDA graph

    def __init__(self, config, dataset, device):
        super(DirectAU, self).__init__(config, dataset, device,
                                       utility.utility_data.data_graph.sparse_adjacency_matrix if config['encoder'] == 'LightGCN' else None)
        self.gamma = float(config['gamma'])

    def fused_trainer(self, lr, max_batch):
        """LightGCN encoder: the fused CUDA-graph step of idgrec.engine (row-restricted propagation, batch x batch loss on
        tensor cores, Adam in the last backward epilogue).  MF encoder: autograd ops + torch.optim.Adam."""
        if self.config['encoder'] == 'MF':
            return None
        return super(DirectAU, self).fused_trainer(lr, max_batch)

    def aggregate(self):
        return self._split(self.encode())

    def forward(self, user, positive, negative):
        """DirectAU.py:59-79 -> [align, gamma * (uniform_u + uniform_i) / 2, reg_lambda * reg(ego_u, ego_pos)]."""
        E0 = self.table()
        final = self.encode(E0)
        ue, pe = self.batch_rows(final, user, positive)
        align_loss = ops.pair_loss("align", ue, pe)
        uniform_loss = self.gamma * (ops.pair_loss("uniform", ue) + ops.pair_loss("uniform", pe)) / 2
        # ego L2 over (user, positive) only: reg_mask 3; the BPR output of the fused kernel is not part of the loss list
        reg_loss = ops.bpr_reg_loss(final.detach(), E0, user, positive, negative, self.dataset.num_users, self.reg_lambda, 3)[1]
        return [align_loss, uniform_loss, reg_loss]


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = DirectAU(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger

    def train(self):
        trainer.universal_trainer(self.model, self.args, self.config, self.dataset, self.device, self.logger)
