"""utility_function/losses.py of the reference (losses.py:4-35): tensor-in / tensor-out,
autograd-capable, computed by the fused CUDA kernels of libidgrec_sm100.so (no torch math).
The models' fused training path bypasses these and calls the same kernels on whole tables."""
import torch

from idgrec import ops


def get_bpr_loss(user_embedding, positive_embedding, negative_embedding):
    """mean(-log(sigmoid(<u,p> - <u,n>) + 1e-7))   (losses.py:4-13)."""
    B = user_embedding.shape[0]
    F = torch.cat([user_embedding, positive_embedding, negative_embedding])
    idx = torch.arange(B, device=F.device)
    return ops.bpr_reg_loss(F, F, idx, idx, idx + B, B, 0.0, 0)[0]


def get_reg_loss(*embeddings):
    """sum_t 0.5*||E_t||^2 / rows(E_t)   (losses.py:16-21)."""
    total = 0
    for e in embeddings:
        n = e.shape[0]
        E = torch.cat([e, e[:1]])  # one dummy "item" row so that N > U
        idx = torch.arange(n, device=e.device)
        zero = torch.zeros(n, dtype=torch.long, device=e.device)
        total = total + ops.bpr_reg_loss(E.detach(), E, idx, zero, zero, n, 1.0, 1)[1]
    return total


def get_InfoNCE_loss(embedding_1, embedding_2, temperature):
    """in-batch InfoNCE with the 1e-5 guard (losses.py:24-35); rows are already gathered."""
    idx = torch.arange(embedding_1.shape[0], device=embedding_1.device)
    return ops.infonce_rows(embedding_1, embedding_2, idx, temperature)
