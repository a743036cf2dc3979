"""Multi-GPU parity inside `pytest -m gpu`: launches tests/multi_gpu_check.py under torchrun on 2, 4 and 8 GPUs of this
box (skipped for the counts the box does not have). The check compares the N-GPU sharded solve with the single-GPU
solve AND with the CPU oracle on the whole problem, and the replicated state across ranks bit for bit."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20)
        return sum(1 for ln in out.stdout.splitlines() if ln.startswith("GPU ")) if out.returncode == 0 else 0
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_solve_matches_single_gpu_and_oracle(world):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (out.stdout + out.stderr)[-3000:]
    assert out.returncode == 0 and "MULTI_GPU_PARITY OK" in out.stdout, tail
