"""SGL (Wu et al., SIGIR'21) -- same class interface as the reference's models/SGL.py:15-198: LightGCN propagation on
the full graph for BPR and on two edge-dropped sub-graphs (re-drawn every epoch) for the in-batch InfoNCE."""
from time import time

import torch

import utility.utility_data.data_graph
import utility.utility_function.tools as tools
import utility.utility_train.batch_test as batch_test
from idgrec import ops
from idgrec.model_base import PropagationModel


class SGL(PropagationModel):
    kind = "SGL"

    def __init__(self, config, dataset, device):
        super(SGL, self).__init__(config, dataset, device, utility.utility_data.data_graph.sparse_adjacency_matrix)
        self.ssl_lambda = float(config['ssl_lambda'])
        self.temperature = float(config['temperature'])

    fused_trainer = None  # three graphs per step: autograd ops over the CUDA kernels, driven by SGL_trainer below

    def _propagate(self, E0, graph):
        K = self.num_layers
        if isinstance(graph, list):
            # 'rw': one sub-graph per layer (SGL.py:48-49); layer mean over K+1 outputs
            x, acc = E0, E0
            for layer in range(K):
                x = ops.spmm(x, graph[layer])
                acc = acc + x
            return acc / float(K + 1)
        return ops.propagate(E0, graph, K, include_layer0=True)

    def aggregate(self, graph=None):
        """SGL.py:39-58 (``graph`` defaults to the full adjacency)."""
        return self._split(self._propagate(self.table(), self.Graph if graph is None else graph))

    def forward(self, user, positive, negative, sub_graph_1, sub_graph_2):
        """SGL.py:60-89 -> [bpr, reg_lambda * reg, ssl_lambda * (user InfoNCE + item InfoNCE)].  The InfoNCE rows are the
        batch's users / positives WITH duplicates (SGL.py:80-84), gathered into dense blocks first."""
        U = self.dataset.num_users
        E0 = self.table()
        final = self._propagate(E0, self.Graph)
        v1 = self._propagate(E0, sub_graph_1)
        v2 = self._propagate(E0, sub_graph_2)
        loss = ops.bpr_reg_loss(final, E0, user, positive, negative, U, self.reg_lambda, 7)
        rows = torch.arange(user.numel(), device=E0.device)
        uidx, iidx = user.long(), positive.long() + U
        ssl = ops.infonce_rows(ops.gather_rows(v1, uidx), ops.gather_rows(v2, uidx), rows, self.temperature) \
            + ops.infonce_rows(ops.gather_rows(v1, iidx), ops.gather_rows(v2, iidx), rows, self.temperature)
        return [loss[0], loss[1], self.ssl_lambda * ssl]

    def get_rating_for_test(self, user):
        with torch.no_grad():
            users_emb, items_emb = self.aggregate(self.Graph)
            return ops.rating_matrix(users_emb, items_emb, user)


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = SGL(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger
        self.aug_type = config['aug_type']
        self.ssl_ratio = float(config['ssl_ratio'])

    def train(self):
        self.SGL_trainer()

    def draw_sub_graphs(self):
        """SGL.py:134-147: two sub-graphs per epoch ('ed'/'nd'), or two lists of per-layer sub-graphs ('rw')."""
        def one():
            return tools.convert_sp_mat_to_sp_tensor(tools.create_adj_mat(self.dataset.user_item_net, self.aug_type, self.ssl_ratio)).to(self.device)
        if self.aug_type in ['nd', 'ed']:
            g1 = one()
            g2 = one()
            return g1, g2
        g1, g2 = [], []
        for _ in range(0, int(self.config['GCN_layer'])):
            g1.append(one())
            g2.append(one())
        return g1, g2

    def SGL_trainer(self):
        """SGL.py:119-198: same epoch structure, log lines and final test as the reference."""
        self.model.to(self.device)
        Optim = torch.optim.Adam(self.model.parameters(), lr=float(self.config['learn_rate']))
        best_results = dict()
        best_results['count'] = 0
        best_results['epoch'] = 0
        best_results['recall'] = [0. for _ in eval(self.config['top_K'])]
        best_results['ndcg'] = [0. for _ in eval(self.config['top_K'])]
        batch_size = int(self.config['batch_size'])
        import utility.utility_train.trainer as trainer

        for epoch in range(int(self.config['training_epochs'])):
            print('-' * 100)
            start_time = time()
            sub_graph_1, sub_graph_2 = self.draw_sub_graphs()
            self.model.train()
            users, pos_items, neg_items = trainer.sample_epoch(self.dataset, torch.device(self.device))
            num_batch = len(users) // batch_size + 1
            total = None
            for batch_users, batch_positive, batch_negative in tools.mini_batch(users, pos_items, neg_items, batch_size=batch_size):
                loss_list = self.model(batch_users, batch_positive, batch_negative, sub_graph_1, sub_graph_2)
                assert len(loss_list) >= 1
                stacked = torch.stack([l.reshape(()) for l in loss_list])
                Optim.zero_grad()
                stacked.sum().backward()
                Optim.step()
                total = stacked.detach() if total is None else total + stacked.detach()   # one device read per epoch
            total_loss_list = total.cpu().tolist()
            end_time = time()
            loss_strs = str(round(sum(total_loss_list) / num_batch, 6)) \
                + " = " + " + ".join([str(round(i / num_batch, 6)) for i in total_loss_list])
            print("\t Epoch: %4d| train time: %.3f | train_loss: %s" % (epoch + 1, end_time - start_time, loss_strs))
            self.logger.info("Epoch: %4d | Training time: %.3f | training loss: %s" % (epoch + 1, end_time - start_time, loss_strs))
            if epoch % int(self.config['interval']) == 0:
                result, best_results = batch_test.general_test(self.dataset, self.model, self.device, self.config, epoch, best_results)
                self.logger.info("Epoch: %4d | Test recall: %s | Test NDCG: %s" % (epoch + 1, result['recall'], result['ndcg']))

        print("\t Model training process completed.")
        self.logger.info('Model training process completed.')
        result, best_results = batch_test.general_test(self.dataset, self.model, self.device, self.config,
                                                       int(self.config['training_epochs']), best_results)
        self.logger.info("Best epoch: %4d | Best recall: %s | Best NDCG: %s" % (best_results['epoch'], best_results['recall'], best_results['ndcg']))
