"""utility_function/metrics.py of the reference (metrics.py:4-58), vectorised numpy, float64.
The evaluator computes the same sums on the device (idg_eval_metrics); these host versions keep
the functional API for callers that hold a hit matrix."""
import numpy as np


def get_label(true_data, pred_data):
    """metrics.py:49-58: r[i, j] = pred_data[i][j] in true_data[i]."""
    r = np.zeros((len(true_data), len(pred_data[0]) if len(pred_data) else 0), dtype="float")
    for i, truth in enumerate(true_data):
        r[i] = np.isin(np.asarray(pred_data[i]), np.asarray(list(truth)))
    return r


def recall_at_k(r, k, test_data):
    hits = r[:, :k].sum(1)
    n = np.array([len(t) for t in test_data])
    return np.sum(hits / n)


def precision_at_k(r, k, test_data):
    return np.sum(r[:, :k].sum(1)) / k


def ndcg_at_k(r, k, test_data):
    assert len(r) == len(test_data)
    disc = 1. / np.log2(np.arange(2, k + 2))
    n = np.array([min(k, len(t)) for t in test_data])
    ideal = (np.arange(k)[None, :] < n[:, None]).astype(float)
    idcg = np.sum(ideal * disc, axis=1)
    dcg = np.sum(r[:, :k] * disc, axis=1)
    idcg[idcg == 0.] = 1.
    ndcg = dcg / idcg
    ndcg[np.isnan(ndcg)] = 0.
    return np.sum(ndcg)
