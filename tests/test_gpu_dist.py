"""Multi-GPU parity (needs >= 2 GPUs): the row-partitioned trainer reproduces the 1-GPU trainer bit for bit."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import REPO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("graph,closure", [("1", "auto"), ("0", "auto"), ("1", "1")])
def test_two_gpu_training_bit_identical(graph, closure):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, IDG_GRAPH=graph, IDG_CLOSURE=closure)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", os.path.join(REPO, "tools", "dist_check.py"), "small", "4"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert "DIST_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("kind", ["SimGCL", "XSimGCL", "XSimGCL3"])
def test_two_gpu_contrastive_training_bit_identical(kind):
    """Row-partitioned SimGCL / XSimGCL (cl_layer = 1 and = K) with injected noise: losses and tables bit-identical to
    the single-GPU fused trainer, sharded evaluation identical (tools/dist_check.py)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29741", os.path.join(REPO, "tools", "dist_check.py"), "small", "3", kind],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert "DIST_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("kind,graph", [("LightGCN", "1"), ("LightGCN", "0"), ("SimGCL", "1"), ("XSimGCL3", "1")])
def test_two_gpu_chunked_exchange_bit_identical(kind, graph):
    """The chunked exchange (idgrec/dist.py: local rows in nnz-balanced blocks, finished blocks streamed to the peers by a few
    CTAs on a second stream -- the default from 8 GPUs up) gives the same bits as the in-epilogue peer stores and as one GPU."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, IDG_GRAPH=graph, IDG_DIST_EXCHANGE="chunked", IDG_DIST_CHUNKS="3", IDG_PUSH_CTAS="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29751", os.path.join(REPO, "tools", "dist_check.py"), "small", "3", kind],
                       capture_output=True, text=True, timeout=600, env=env)
    assert "DIST_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
