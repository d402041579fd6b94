"""NGCF (Wang et al., SIGIR'19) -- same class interface as the reference's models/NGCF.py:17-153.
Graph D^-1/2 (A+I) D^-1/2; per layer side = A.E (hand-written SpMM kernel), then one fused
dense kernel per layer: sum = side W_gcn + b, bi = (E * side) W_bi + b, LeakyReLU(0.2), message dropout,
row-normalise (csrc/ngcf.cu, forward and backward); final = concat of the K+1 blocks (256-d).  BPR / regulariser /
full-ranking evaluation run on the fused kernels (d = 256 for scores)."""
import torch
from torch import nn

import utility.utility_data.data_graph
import utility.utility_train.trainer as trainer
from idgrec import ops
from idgrec.model_base import PropagationModel


class NGCF(PropagationModel):
    kind = "NGCF"
    graph_capturable = True   # forward() has no host sync: universal_trainer replays the whole step from a CUDA graph

    def fused_trainer(self, lr, max_batch):
        """The whole step (propagation, dense layers, BPR, both backward chains, Adam) without autograd, replayed from a
        CUDA graph (idgrec/engine_ngcf.py).  ``fused_step = 0`` in the config keeps the autograd ops + torch.optim.Adam."""
        if str(self.config.get('fused_step', '1')) in ('0', 'False', 'false'):
            return None
        if self._fused is None or self._fused.lr != lr or self._fused.max_batch < max_batch:
            from idgrec.engine_ngcf import NgcfFusedTrainer
            self._fused = NgcfFusedTrainer(self, lr, max_batch, use_cuda_graph=str(self.config.get('cuda_graph', '1')) not in ('0', 'False', 'false'))
        return self._fused

    def __init__(self, config, dataset, device):
        super(NGCF, self).__init__(config, dataset, device, None)
        initializer = nn.init.xavier_uniform_
        self.weight_dict = nn.ParameterDict()
        layers = [int(config['embedding_size'])] + eval(config['layer_size'])
        # same parameter order (and torch RNG draws) as NGCF.py:36-44
        for layer in range(int(config['GCN_layer'])):
            self.weight_dict.update({'W_gcn_%d' % layer: nn.Parameter(initializer(torch.empty(layers[layer], layers[layer + 1])))})
            self.weight_dict.update({'b_gcn_%d' % layer: nn.Parameter(initializer(torch.empty(1, layers[layer + 1])))})
            self.weight_dict.update({'W_bi_%d' % layer: nn.Parameter(initializer(torch.empty(layers[layer], layers[layer + 1])))})
            self.weight_dict.update({'b_bi_%d' % layer: nn.Parameter(initializer(torch.empty(1, layers[layer + 1])))})
        if eval(config['mess_dropout']):
            self.mess_dropout = eval(config['mess_drop_prob'])
        if eval(config['node_dropout']):
            raise NotImplementedError("node_dropout = True is not part of the accelerated path (the reference's default is False)")
        import utility.utility_function.tools as tools
        self.Graph = utility.utility_data.data_graph.sparse_adjacency_matrix_with_self(dataset)
        self.Graph = tools.convert_sp_mat_to_sp_tensor(self.Graph)
        self.Graph = self.Graph.coalesce().to(self.device)
        self.activation_layer = nn.Tanh()

    def aggregate(self, keep_masks=None):
        """NGCF.py:67-111.  ``keep_masks`` (K tensors of 0/1) injects the dropout draws for parity tests;
        otherwise nn.Dropout is constructed inline exactly like the reference -- i.e. ALWAYS active, also in
        eval mode."""
        ego = self.table()
        outs = [ego]
        wd = self.weight_dict
        for layer in range(int(self.config['GCN_layer'])):
            p = float(self.mess_dropout[layer])
            if keep_masks is not None:
                keep = keep_masks[layer]
            else:  # the draw nn.Dropout(p) makes (Bernoulli(1-p) per element, torch device generator), always on
                keep = (torch.rand_like(ego.detach()) >= p).float()
            # SpMM + fused dense epilogue (two 64x64 products, biases, E*side, LeakyReLU, dropout, row-normalise)
            ego, norm = ops.ngcf_layer(ego, wd['W_gcn_%d' % layer], wd['b_gcn_%d' % layer], wd['W_bi_%d' % layer], wd['b_bi_%d' % layer],
                                       self.Graph, keep, p)
            outs.append(norm)
        final = torch.cat(outs, dim=1)
        return self._split(final)

    def forward(self, user, positive, negative, keep_masks=None):
        """NGCF.py:113-130 -> [bpr, reg]; the regulariser covers the item ego rows only (NGCF.py:120-125)."""
        U = self.dataset.num_users
        users_emb, items_emb = self.aggregate(keep_masks)
        final = torch.cat([users_emb, items_emb])
        E0 = self.table()
        bpr = ops.bpr_reg_loss(final, final, user, positive, negative, U, 0.0, 0)[0]
        reg = ops.bpr_reg_loss(E0.detach(), E0, user, positive, negative, U, self.reg_lambda, 6)[1]
        return [bpr, reg]


class Trainer():
    def __init__(self, args, config, dataset, device, logger):
        self.model = NGCF(config, dataset, device)
        self.args, self.device, self.config, self.dataset, self.logger = args, device, config, dataset, logger

    def train(self):
        trainer.universal_trainer(self.model, self.args, self.config, self.dataset, self.device, self.logger)
