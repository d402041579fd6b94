"""Matrix factorisation with BPR (models/MFBPR.py:12-64 of the reference): no graph; the same fused
BPR and top-K kernels, useful as the smallest end-to-end model."""
import utility.utility_train.trainer as trainer
from idgrec import ops
from idgrec.model_base import PropagationModel


class MFBPR(PropagationModel):
    kind = "MFBPR"

    def __init__(self, config, dataset, device):
        super(MFBPR, self).__init__(config, dataset, device, None)

    def aggregate(self):
        return self._split(self.table())

    def forward(self, user, positive, negative):
        E0 = self.table()
        loss = ops.bpr_reg_loss(E0, E0, user, positive, negative, self.dataset.num_users, self.reg_lambda, 7)
        return [loss[0], loss[1]]


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = MFBPR(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger

    def train(self):
        trainer.universal_trainer(self.model, self.args, self.config, self.dataset, self.device, self.logger)
