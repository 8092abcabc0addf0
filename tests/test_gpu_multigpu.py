"""Multi-GPU parity (SURVEY.md section 4 layer 4), skipped on boxes with fewer than 2 GPUs: torchrun launches
tests/multigpu_worker.py on every visible GPU (at most 8); see that file for what is checked."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (run with gpurun --gpus 2)")
def test_sharded_step_and_nvls_allreduce_across_gpus():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(REPO, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-4000:]); sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0, "multigpu_worker failed:\n" + r.stdout[-3000:] + r.stderr[-3000:]
    assert "ok=1" in r.stdout
