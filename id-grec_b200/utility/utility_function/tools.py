"""utility_function/tools.py of the reference, hot-path subset (tools.py:8-64,95-109).

Same names and argument meaning.  ``shuffle``/``mini_batch``/``set_seed`` keep the numpy
global-RNG stream the reference consumes so epochs >= 1 see bit-identical batches.
"""
import os
import random

import numpy as np
import torch


def set_seed(seed):
    """tools.py:8-14."""
    np.random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.manual_seed(seed)


def read_configuration(filename, model):
    """tools.py:17-32: ``key = value`` lines into a dict of strings."""
    if not os.path.exists(filename):
        print("\tThe path does not have a configuration file for " + model + ".")
        raise IOError
    config = dict()
    with open(filename, "r") as f:
        for line in f:
            if line == "":
                break
            if line.lstrip().startswith("#") or not line.strip():
                continue          # comment / blank lines (the shipped files carry a provenance header)
            try:
                name, value = line.strip().split("=")
                config[name.strip()] = value.strip()
            except ValueError:
                print("\tConfiguration file format error.")
    return config


def shuffle(*arrays, **kwargs):
    """tools.py:35-52: one np.random.shuffle of arange(n), applied to every array."""
    require_indices = kwargs.get('indices', False)
    if len(set(len(x) for x in arrays)) != 1:
        raise ValueError('Inputs to shuffle must have the same length.')
    shuffle_indices = np.arange(len(arrays[0]))
    np.random.shuffle(shuffle_indices)
    if torch.is_tensor(arrays[0]) and arrays[0].is_cuda:
        sel = torch.from_numpy(shuffle_indices).to(arrays[0].device)
    else:
        sel = shuffle_indices
    result = arrays[0][sel] if len(arrays) == 1 else tuple(x[sel] for x in arrays)
    return (result, shuffle_indices) if require_indices else result


def mini_batch(*tensors, **kwargs):
    """tools.py:55-64: consecutive slices, last one short."""
    batch_size = kwargs.get('batch_size', 1024)
    if len(tensors) == 1:
        tensor = tensors[0]
        for i in range(0, len(tensor), batch_size):
            yield tensor[i:i + batch_size]
    else:
        for i in range(0, len(tensors[0]), batch_size):
            yield tuple(x[i:i + batch_size] for x in tensors)


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
