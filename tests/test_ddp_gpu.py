"""Two-GPU parity of the data-parallel training step (NCCL): the gradient arena after the bucketed, overlapped
all-reduce equals the SUM of the per-shard oracle gradients, per parameter, at full ResNet-50 depth;
and the overlapped reduction equals a plain sum with identical parameters on every rank afterwards.  Skipped on boxes
with fewer than two GPUs (run with `gpurun --gpus 2`; log committed under profiles/)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, timeout=1500):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "ddp_check.py")] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_gradients_equal_summed_per_shard_oracle():
    r = _torchrun(["tdo", "--oracle", "--full"])
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "oracle step 1" in r.stdout


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_bucketed_update_matches_plain_path():
    """Overlapped reduce + per-bucket optimizer update (FusedTrainer.step) == plain all-reduce + one update."""
    r = _torchrun(["tdo", "--step"])
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "fused step 2" in r.stdout
