"""Hardware multi-GPU parity (SURVEY.md §8(e) option (i)): needs >= 2 visible GPUs (run with `gpurun --gpus 2`), skipped otherwise.
The gloo world-2 tests in test_distributed_cpu.py cover the same host logic on CPU."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_nccl_allreduced_gradient_equals_mean_of_per_shard_gradients():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_mp_grad_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("MULTI_GPU_PARITY")]
    assert r.returncode == 0 and line and line[0].endswith("ok=True"), (r.stdout[-1500:], r.stderr[-1500:])
    print(line[0])
