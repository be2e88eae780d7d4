"""GPU, needs >= 2 devices (skipped otherwise): launches tests/run_slab_gpu.py under torchrun -- slab
decomposition with the peer-memory halo exchange (boundary kernel stores into the neighbour's ghost rows
over NVLink) and with NCCL send/recv, the domain-divided mod_main + rtm_main shot, and the shot-parallel
chained image stack, each bit for bit against one GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_gpu_slab_and_shot_partitioning_bitwise():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs (one process per GPU)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "run_slab_gpu.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "MISMATCH" not in out
    for what in ("halo=p2p vs single domain bitwise: OK", "halo=nccl vs single domain bitwise: OK",
                 "mod_main+rtm_main x2 halo=p2p vs one GPU bitwise: OK", "chained stack vs sequential bitwise: OK"):
        assert what in out, out[-4000:]
