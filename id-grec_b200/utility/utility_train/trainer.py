"""utility_train/trainer.py of the reference (trainer.py:8-74): same entry point, same epoch
structure (sample -> shuffle -> mini-batches -> step -> periodic full-ranking test -> early stop),
same log lines.  Models that expose ``fused_trainer`` run each step as fused CUDA kernels
(propagate -> loss -> backward -> Adam, replayed from a CUDA graph, losses accumulated on the
device and read once per epoch instead of two ``.item()`` syncs per batch, trainer.py:52); any
other nn.Module goes through the reference's autograd + torch.optim.Adam loop unchanged."""
from time import time

import numpy as np
import torch

import utility.utility_function.tools as tools
import utility.utility_train.batch_test as batch_test


def sample_epoch(dataset, device):
    """trainer.py:26-35: sample, move to the device, shuffle with the numpy global stream.
    ids stay int64 end to end (the reference's float32 round-trip is exact below 2^24).  On CUDA only the
    negatives and the permutation cross PCIe (one pinned [2, E] int64 block): the train edges are resident on the
    device and tools.shuffle's fancy indexing runs there (idg_permute3)."""
    device = torch.device(device)
    if device.type != "cuda":
        sample_data = dataset.sample_data_to_train_all()
        perm = np.arange(len(sample_data))
        np.random.shuffle(perm)                   # tools.shuffle (tools.py:41-42), same stream position
        t = torch.from_numpy(np.ascontiguousarray(sample_data[perm].T))
        return t[0], t[1], t[2]
    return _samples_to_device(dataset, device, _sample_epoch_host(dataset))


def _sample_epoch_host(dataset):
    """Host half of sample_epoch: the epoch's negatives (exact replay of data_loader.py:108-127) and the shuffle permutation
    (tools.py:41-42), drawn from the numpy global stream in the reference's order, into the pinned [2, E] staging block."""
    E = len(dataset.train_user)
    stage = _PINNED.get(E)
    if stage is None:
        stage = _PINNED[E] = torch.empty((2, max(E, 1)), dtype=torch.int64).pin_memory()
    host = stage.numpy()
    host[0, :E] = dataset.sample_negatives()
    perm = host[1, :E]
    perm[:] = np.arange(E)
    np.random.shuffle(perm)                       # same generator position as the reference: right after the sampler
    return stage


def _samples_to_device(dataset, device, stage):
    """Device half: one pinned H2D copy, then tools.shuffle's fancy indexing on the device (idg_permute3)."""
    from idgrec import _lib
    E = len(dataset.train_user)
    cache = dataset.device_cache(device)
    t = stage.to(device, non_blocking=True)
    _check_same_samples(t, E)
    out = torch.empty((3, E), dtype=torch.int64, device=device)
    _lib.check(_lib.lib().idg_permute3(_lib.ptr(cache["train_user"]), _lib.ptr(cache["train_item"]), _lib.ptr(t[0]), _lib.ptr(t[1]), E,
                                       _lib.ptr(out), _lib.cur_stream()), "idg_permute3")
    torch.cuda.current_stream().synchronize()      # the staging buffer is reused by the next epoch's prefetch
    return out[0], out[1], out[2]


class EpochPrefetch:
    """The next epoch's host sampling on a worker thread, started BEFORE the current epoch's steps are enqueued: the launch queue
    holds only a few hundred steps, so the enqueue loop of an epoch runs for as long as the GPU does and sampling after it
    (round 1) overlapped only its tail.  numpy's generators and the C sampler walk release the GIL; nothing else reads the numpy
    global stream while the worker runs, so the stream -- and every batch -- stays the reference's."""

    def __init__(self, dataset):
        import threading
        self.dataset, self.stage, self.error = dataset, None, None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            self.stage = _sample_epoch_host(self.dataset)
        except BaseException as e:  # noqa: BLE001 -- re-raised on the caller's thread
            self.error = e

    def finish(self, device):
        self.thread.join()
        if self.error is not None:
            raise self.error
        return _samples_to_device(self.dataset, device, self.stage)


_PINNED = {}


def _check_same_samples(t, E):
    """Row-partitioned training evaluates the loss redundantly on every rank and never reduces gradients: it is only
    correct if all ranks drew the same negatives and the same permutation (same seed, same numpy stream position).
    One int64 checksum per epoch, min/max all-reduced; a mismatch raises instead of silently corrupting the tables."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    w = torch.arange(1, E + 1, device=t.device, dtype=torch.int64)
    c = ((t[0, :E] * 1000003 + t[1, :E]) * w).sum().reshape(1)
    lo, hi = c.clone(), c.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if int(lo.item()) != int(hi.item()):
        raise RuntimeError("ranks drew different epoch samples (negatives / shuffle differ): seed every rank identically "
                           "(tools.set_seed) and keep other np.random calls off the global stream")


def universal_trainer(model, args, config, dataset, device, logger):
    model.to(device)
    lr = float(config['learn_rate'])
    batch_size = int(config['batch_size'])
    fused = model.fused_trainer(lr, batch_size) if getattr(model, "fused_trainer", None) is not None else None
    if (fused is None and getattr(model, "graph_capturable", False) and torch.device(device).type == "cuda"
            and str(config.get('cuda_graph', '1')) not in ('0', 'False', 'false')):
        # autograd-path models without host syncs in forward(): the whole step (forward, backward, Adam) replays from
        # a CUDA graph; same kernels, same order, same results as the eager loop below
        from idgrec.graphed import GraphedStep
        fused = GraphedStep(model, lr, batch_size)
    Optim = None if fused is not None else torch.optim.Adam(model.parameters(), lr=lr)

    best_results = dict()
    best_results['count'] = 0
    best_results['epoch'] = 0
    best_results['recall'] = [0. for _ in eval(config['top_K'])]
    best_results['ndcg'] = [0. for _ in eval(config['top_K'])]
    best_results['stop'] = 0

    n_epochs = int(config['training_epochs'])
    prefetched = None
    for epoch in range(n_epochs):
        print('-' * 100)
        start_time = time()
        model.train()

        users, pos_items, neg_items = prefetched if prefetched is not None else sample_epoch(dataset, device)
        prefetched = None
        num_batch = len(users) // batch_size + 1          # trainer.py:36 (over-counts when divisible)

        if fused is not None:
            # draw the next epoch's samples on a worker thread while this epoch's steps are enqueued and run.  Nothing else
            # reads the numpy global generator in between, so the stream (and every batch) stays the reference's.
            worker = EpochPrefetch(dataset) if (epoch + 1 < n_epochs and torch.device(device).type == "cuda") else None
            for bu, bp, bn in tools.mini_batch(users, pos_items, neg_items, batch_size=batch_size):
                fused.step(bu, bp, bn)
            enqueue_time = time()
            if worker is not None:
                prefetched = worker.finish(device)
            elif epoch + 1 < n_epochs:
                prefetched = sample_epoch(dataset, device)
            sampling_time = time() - enqueue_time         # what the NEXT epoch's samples still cost after this epoch's steps were enqueued
            total_loss_list = fused.pop_epoch_losses()    # one device read per epoch (waits for the epoch's last step)
        else:
            acc = None
            for bu, bp, bn in tools.mini_batch(users, pos_items, neg_items, batch_size=batch_size):
                loss_list = model(bu, bp, bn)
                assert len(loss_list) >= 1
                stacked = torch.stack([l.reshape(()) for l in loss_list])
                Optim.zero_grad()
                stacked.sum().backward()
                Optim.step()
                # per-loss epoch sums stay on the device: one read per epoch instead of len(loss_list) .item() syncs
                # per batch (trainer.py:52)
                acc = stacked.detach().double() if acc is None else acc + stacked.detach()
            total_loss_list = acc.cpu().tolist() if acc is not None else []
            sampling_time = 0.0

        # "Training time" = this epoch's steps: the overlapped sampling of the next epoch only counts where it outlasted them
        end_time = time()
        loss_strs = str(round(sum(total_loss_list) / num_batch, 6)) \
            + " = " + " + ".join([str(round(i / num_batch, 6)) for i in total_loss_list])
        print("Training time: %.3f | training loss: %s" % (end_time - start_time, loss_strs))
        if sampling_time > 0.0:
            print("\t(next epoch's host sampling: %.3f s, overlapped with this epoch's kernels)" % sampling_time)
        logger.info("Epoch: %4d | Training time: %.3f | training loss: %s" % (epoch + 1, end_time - start_time, loss_strs))

        if epoch % int(config.get('interval', 1)) == 0:   # configure/DirectAU.txt has no interval line
            result, best_results = batch_test.general_test(dataset, model, device, config, epoch, best_results)
            logger.info("Epoch: %4d | Test recall: %s | Test NDCG: %s" % (epoch + 1, result['recall'], result['ndcg']))
            if best_results['stop'] > 0:
                break

    print("Model training process completed.")
    logger.info('Model training process completed.')
    logger.info("Best epoch: %4d | Best recall: %s | Best NDCG: %s" % (best_results['epoch'], best_results['recall'], best_results['ndcg']))
