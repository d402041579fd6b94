"""Whole-step CUDA-graph replay for the models that train through the autograd ops (NGCF, LightCCF, LightCSCF,
DirectAU): forward -> backward -> torch.optim.Adam(capturable) is captured once per batch size and replayed, so
a step costs one graph launch instead of ~100 Python-dispatched kernel launches, and the per-loss epoch sums stay on
the device (the reference reads every loss of every batch with .item(), trainer.py:52).

The arithmetic is exactly the eager loop's (trainer.py:40-56): the same kernels in the same order.  The warm-up
iterations torch needs before a capture are undone (parameters, Adam state, loss sums and the device RNG state are
restored), so the training trajectory does not depend on whether or when a graph was captured."""
from __future__ import annotations

import torch


class GraphedStep:
    def __init__(self, model, lr, max_batch, warmup=2):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.dev = self.params[0].device
        if self.dev.type != "cuda":
            raise RuntimeError("GraphedStep needs the model on a CUDA device; there is no CPU fallback")
        self.lr, self.max_batch, self.warmup = lr, max_batch, warmup
        self.opt = torch.optim.Adam(self.params, lr=lr, capturable=True)
        self.batch = torch.zeros(3, max_batch, dtype=torch.int64, device=self.dev)
        self.loss_acc = None
        self.graphs = {}
        self.replays = 0

    # one training step on the static batch views (eager during warm-up, recorded during capture)
    def _one(self, views):
        loss_list = self.model(*views)
        assert len(loss_list) >= 1
        stacked = torch.stack([l.reshape(()) for l in loss_list])
        self.opt.zero_grad(set_to_none=True)
        stacked.sum().backward()
        self.opt.step()
        if self.loss_acc is None:
            self.loss_acc = torch.zeros_like(stacked.detach(), dtype=torch.float64)   # Python-float sums in the reference (trainer.py:52-53)
        self.loss_acc.add_(stacked.detach())

    def _snapshot(self):
        st = []
        for p in self.params:
            s = self.opt.state.get(p, {})
            st.append({k: v.detach().clone() for k, v in s.items() if torch.is_tensor(v)})
        return ([p.detach().clone() for p in self.params], st, None if self.loss_acc is None else self.loss_acc.clone(),
                torch.cuda.get_rng_state(self.dev))

    def _restore(self, snap):
        saved_p, saved_s, saved_acc, rng = snap
        with torch.no_grad():
            for p, q in zip(self.params, saved_p):
                p.copy_(q)
            for p, old in zip(self.params, saved_s):
                for k, v in self.opt.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        v.copy_(old[k]) if k in old else v.zero_()
            if self.loss_acc is not None:
                self.loss_acc.copy_(saved_acc) if saved_acc is not None else self.loss_acc.zero_()
        torch.cuda.set_rng_state(rng, self.dev)

    def _capture(self, B):
        views = tuple(self.batch[i, :B] for i in range(3))
        snap = self._snapshot()
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._one(views)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g):
            self._one(views)
        self._restore(snap)
        self.graphs[B] = g
        return g

    def step(self, user, positive, negative):
        B = int(user.numel())
        if B > self.max_batch:
            raise ValueError("batch of %d exceeds max_batch %d" % (B, self.max_batch))
        self.batch[0, :B].copy_(user)
        self.batch[1, :B].copy_(positive)
        self.batch[2, :B].copy_(negative)
        g = self.graphs.get(B)
        if g is None:
            g = self._capture(B)
        g.replay()
        self.replays += 1

    def pop_epoch_losses(self):
        """Per-loss sums since the last call (one device read)."""
        if self.loss_acc is None:
            return []
        out = self.loss_acc.cpu().tolist()
        self.loss_acc.zero_()
        return out
