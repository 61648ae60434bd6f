"""N > 1 on real GPUs: world_size-2 (and, when the box has them, 4 / 8) torchrun jobs over NCCL -- reads sharded at
newline boundaries, genome replicated, genome-sharded spectrum, tables all-reduced -- must equal the oracle and the
single-GPU result (pss-bam.c:401-420, fragkon.c:137-146, genome-kmer-count.c:56-64).  Skipped on a one-GPU box; the
CPU suite covers the same host logic over gloo (tests/test_multi_rank.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_sharded_equals_single_gpu_and_oracle(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() % 1000) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout and "spectrum k=12 == oracle: ok" in r.stdout
