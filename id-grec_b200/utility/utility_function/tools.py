"""utility_function/tools.py of the reference, hot-path subset (tools.py:8-64,95-109).

Same names and argument meaning.  ``shuffle``/``mini_batch``/``set_seed`` keep the numpy
global-RNG stream the reference consumes so epochs >= 1 see bit-identical batches.
"""
import os
import random

import numpy as np
import torch


def set_seed(seed):
    """tools.py:8-14: seeds the three generators the path draws from -- numpy's legacy global stream (negative
    sampler + epoch shuffle), torch's CPU generator (xavier init) and every CUDA generator (SimGCL noise, dropout)."""
    seed = int(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)           # CPU generator; torch also forwards the seed to the CUDA generators
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def _parse_config_line(text):
    """``key = value`` -> (key, value); None for blank / comment lines; ValueError for anything else."""
    body = text.strip()
    if not body or body.startswith("#"):
        return None
    key, sep, value = body.partition("=")
    if not sep or "=" in value:        # the reference's ``a, b = line.split("=")`` accepts exactly one '='
        raise ValueError(body)
    return key.strip(), value.strip()


def read_configuration(filename, model):
    """tools.py:17-32: the ``key = value`` file of a model as a dict of strings.  A missing file raises IOError;
    malformed lines are reported and skipped, like the reference does."""
    if not os.path.isfile(filename):
        print("\tNo configuration file for model %s at %s." % (model, filename))
        raise IOError(filename)
    config = {}
    with open(filename, "r") as handle:
        for number, text in enumerate(handle, 1):
            try:
                item = _parse_config_line(text)
            except ValueError:
                print("\tConfiguration file format error (line %d of %s)." % (number, filename))
                continue
            if item is not None:
                config[item[0]] = item[1]
    return config


def shuffle(*arrays, **kwargs):
    """tools.py:35-52: ONE ``np.random.shuffle`` of ``arange(n)`` (the numpy global stream, right after the sampler)
    applied to every input; ``indices=True`` also returns the permutation.  Device tensors are permuted on the
    device."""
    lengths = {len(a) for a in arrays}
    if len(lengths) != 1:
        raise ValueError("shuffle: all inputs need the same length, got %s" % sorted(lengths))
    perm = np.arange(lengths.pop())
    np.random.shuffle(perm)
    first = arrays[0]
    sel = torch.from_numpy(perm).to(first.device) if (torch.is_tensor(first) and first.is_cuda) else perm
    out = tuple(a[sel] for a in arrays)
    out = out[0] if len(out) == 1 else out
    return (out, perm) if kwargs.get('indices', False) else out


def mini_batch(*tensors, **kwargs):
    """tools.py:55-64: consecutive ``batch_size`` slices of the inputs (default 1024), the last one short; one input
    yields slices, several yield tuples."""
    step = int(kwargs.get('batch_size', 1024))
    for lo in range(0, len(tensors[0]), step):
        piece = tuple(t[lo:lo + step] for t in tensors)
        yield piece[0] if len(piece) == 1 else piece


def create_adj_mat(inter_graph, aug_type, ssl_rate):
    """tools.py:67-92: the edge-dropped graph of SGL.  The kept edges are drawn exactly like the reference does
    (python ``random.sample`` over the row-major nonzeros of user_item_net -- a stream set_seed() does not seed);
    the symmetric normalisation of the kept edges runs in the device CSR builder (float32 degrees, ``inf -> 0``)."""
    from utility.utility_data.data_graph import EdgeListAdjacency
    num_users, num_items = inter_graph.get_shape()
    user_index, item_index = inter_graph.nonzero()
    if aug_type == 'nd':
        raise NotImplementedError("The method does not implemented.")
    elif aug_type in ['ed', 'rw']:
        edge_number = inter_graph.count_nonzero()
        keep_index = random.sample(range(edge_number), k=int((1 - ssl_rate) * edge_number))
        user_index = np.array(user_index)[keep_index]
        item_index = np.array(item_index)[keep_index]
    else:
        raise ValueError("unknown aug_type %r" % (aug_type,))
    return EdgeListAdjacency(user_index, item_index, num_users, num_items)


def convert_sp_mat_to_sp_tensor(sp_mat):
    """tools.py:95-109.  The reference turns a scipy matrix into a torch COO tensor here; the
    adjacency built by utility_data.data_graph already lives on the device as CSR, so it passes
    through unchanged and ``.coalesce().to(device)`` (models/LightGCN.py:32) yields the
    propagation handle."""
    from utility.utility_data.data_graph import NormAdjacency
    if isinstance(sp_mat, NormAdjacency):
        return sp_mat
    raise TypeError("convert_sp_mat_to_sp_tensor expects the adjacency returned by utility_data.data_graph "
                    "(there is no torch.sparse path in this implementation)")
