"""The NVLink peer-memory all-reduce (csrc/peer.cu: one-shot for 2 ranks, two-shot reduce-scatter + all-gather for
more) with one process per GPU: sums equal NCCL's to 1e-12, bits identical on every rank, odd counts and counts
not divisible by the world size included (tools/peer_allreduce_check.py does the work and the timing).
Skips itself on a box with fewer than two GPUs -- `gpurun --gpus 2|8 -- python -m pytest tests/test_gpu_peer.py -m gpu`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_peer_allreduce_matches_nccl(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "peer_allreduce_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("peer_allreduce")]
    assert len(lines) >= 5 and all("rank-identical bits yes" in l for l in lines), p.stdout
