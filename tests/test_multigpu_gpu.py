"""Two-rank check of the in-kernel gradient exchange (csrc/ppo_fused.cu, core/parallel.py) — `-m gpu`, skipped on
boxes with fewer than 2 GPUs.  Spawns scripts/multigpu_check.py under torchrun: the persistent kernel's
reduce-scatter + all-gather over peer memory equals the CUDA-graph + NCCL all-reduce path to 2e-5 after 8 minibatch
steps, weights stay bit-identical across ranks over 3 more learn() calls, Ape-X's sharded PER keeps the global
max-weight normalisation; the continuous-policy network (flat buffer not a multiple of 8 floats) takes the same exchange."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_in_kernel_exchange_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, JB_NENV="64", JB_NEPOCH="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "multigpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "in-kernel gradient exchange ON" in r.stdout
    assert r.stdout.count("PPO dp ok") == 2 and r.stdout.count("Ape-X sharded PER ok") == 2
    assert r.stdout.count("PPO continuous dp ok (p2p=True") == 2          # the obs-11 / act-3 network: num_flat % 8 == 4
