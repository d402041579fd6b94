"""One forward+backward call of the LightCCF batch x batch loss at B = 4096 (for `ncu -k regex:pl_`)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "id-grec_b200"))
from idgrec import ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X = torch.randn(n, 64, device=dev).requires_grad_(True)
Y = (torch.randn(n, 64, device=dev) + 0.5 * X.detach()).requires_grad_(True)
for _ in range(2):
    ops.pair_loss("lightccf", X, Y, 0.22).backward()
torch.cuda.synchronize()
