"""Row strips as they run in production: one process per GPU under torchrun, peers reached through CUDA IPC.
Needs at least two GPUs (skipped otherwise; the single-GPU suite covers the same kernels and protocol with all
strips emulated in one process)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_strips_over_ipc_bit_identical_to_one_strip(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(29560 + world), os.path.join(ROOT, "tools", "strips_check.py")],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["strips"] == world
