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
    ids stay int64 end to end (the reference's float32 round-trip is exact below 2^24)."""
    sample_data = dataset.sample_data_to_train_all()
    perm = np.arange(len(sample_data))
    np.random.shuffle(perm)                       # tools.shuffle (tools.py:41-42), same stream position
    E = len(sample_data)
    if device.type != "cuda":
        t = torch.from_numpy(np.ascontiguousarray(sample_data[perm].T))
        return t[0], t[1], t[2]
    # gather straight into a reused pinned staging buffer, one async H2D copy of [3, E] int64
    stage = _PINNED.get(E)
    if stage is None:
        stage = _PINNED[E] = torch.empty((3, E), dtype=torch.int64).pin_memory()
    host = stage.numpy()
    for c in range(3):
        np.take(sample_data[:, c], perm, out=host[c])
    t = stage.to(device, non_blocking=True)
    torch.cuda.current_stream().synchronize()      # the staging buffer is reused by the next epoch's prefetch
    return t[0], t[1], t[2]


_PINNED = {}


def universal_trainer(model, args, config, dataset, device, logger):
    model.to(device)
    lr = float(config['learn_rate'])
    batch_size = int(config['batch_size'])
    fused = model.fused_trainer(lr, batch_size) if getattr(model, "fused_trainer", None) is not None else None
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
            for bu, bp, bn in tools.mini_batch(users, pos_items, neg_items, batch_size=batch_size):
                fused.step(bu, bp, bn)
            # the steps are only enqueued: draw the next epoch's samples on the host while the GPU trains.  Nothing else
            # reads the numpy global generator in between, so the stream (and every batch) stays the reference's.
            if epoch + 1 < n_epochs:
                prefetched = sample_epoch(dataset, device)
            total_loss_list = fused.pop_epoch_losses()    # one device read per epoch
        else:
            total_loss_list = []
            for batch_i, (bu, bp, bn) in enumerate(tools.mini_batch(users, pos_items, neg_items, batch_size=batch_size)):
                loss_list = model(bu, bp, bn)
                if batch_i == 0:
                    assert len(loss_list) >= 1
                    total_loss_list = [0.] * len(loss_list)
                total_loss = 0.
                for i in range(len(loss_list)):
                    total_loss += loss_list[i]
                    total_loss_list[i] += loss_list[i].item()
                Optim.zero_grad()
                total_loss.backward()
                Optim.step()

        end_time = time()
        loss_strs = str(round(sum(total_loss_list) / num_batch, 6)) \
            + " = " + " + ".join([str(round(i / num_batch, 6)) for i in total_loss_list])
        print("Training time: %.3f | training loss: %s" % (end_time - start_time, loss_strs))
        logger.info("Epoch: %4d | Training time: %.3f | training loss: %s" % (epoch + 1, end_time - start_time, loss_strs))

        if epoch % int(config['interval']) == 0:
            result, best_results = batch_test.general_test(dataset, model, device, config, epoch, best_results)
            logger.info("Epoch: %4d | Test recall: %s | Test NDCG: %s" % (epoch + 1, result['recall'], result['ndcg']))
            if best_results['stop'] > 0:
                break

    print("Model training process completed.")
    logger.info('Model training process completed.')
    logger.info("Best epoch: %4d | Best recall: %s | Best NDCG: %s" % (best_results['epoch'], best_results['recall'], best_results['ndcg']))
