"""GPU: the peer-memory multi-GPU schemes -- whole-row peer push (partition.PushPartition,
ndcn_solver_set_peers) and feature-sharded peer push (partition.FeaturePushPartition,
ndcn_solver_set_feature_peers).

The kernels that produce a gather source also store every new row (or column slice) into the other ranks'
buffers and a device barrier kernel orders those stores before the next gather.
  * test_push_ranks_in_one_process: two to eight "ranks" run as threads of ONE process on ONE GPU
    (tests/push_inproc_worker.py) -- the device addresses of the other ranks' workspaces are valid as
    they are, no IPC needed -- so the 1-GPU box exercises the whole path: full-halo graphs, the pushes
    from the tcgen05 stage kernels / the FP32-FMA kernels / the ground-truth dynamics kernels / the
    pre-stage algebra, the slice scatter / z-owner scatter of the feature-sharded variant, the barrier and
    its 2-double all-reduce.  Checked against the same solve on the
    unpartitioned graph (itself parity-tested against the oracle) and, for one case, the CPU oracle.
  * test_push_two_gpus_torchrun: the real configuration, one process per GPU over CUDA IPC / NVLink
    (tests/push_worker.py); needs >= 2 GPUs.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_push_ranks_in_one_process():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "push_inproc_worker.py")],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "PUSH_INPROC_OK" in out.stdout, out.stdout[-4000:] + out.stderr[-4000:]


def test_push_two_gpus_torchrun():
    """One process per GPU, CUDA IPC over NVLink: the 2-rank solves against the single-GPU solves."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "push_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "PUSH_WORKER_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
