"""GPU: launches tests/run_slab_gpu.py under torchrun -- slab decomposition with the peer-memory halo exchange
(boundary kernel stores into the neighbour's ghost rows, in-kernel acquire/release, CUDA-graph level loop), the
domain-divided mod_main + rtm_main shot and the shot-parallel chained image stack, each bit for bit against one
GPU.  Two flavours: two processes on ONE device (always runs: CUDA IPC maps a same-device peer just as well, gloo
is the rendezvous) and one process per GPU with NCCL (needs >= 2 devices)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(port, env_extra, expect):
    env = dict(os.environ)
    env.update(env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "run_slab_gpu.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "MISMATCH" not in out
    for what in expect:
        assert what in out, out[-4000:]
    return out


@pytest.mark.gpu
def test_slab_and_shot_partitioning_bitwise():
    import torch

    expect = ("halo=p2p vs single domain bitwise: OK", "mod_main+rtm_main x2 halo=p2p vs one GPU bitwise: OK",
              "chained stack vs sequential bitwise: OK", "device stack chained vs sequential bitwise: OK")

    def counter(out, what):
        line = [ln for ln in out.splitlines() if ln.startswith(what)][0]
        return int(line.split(":")[1])

    # thin slabs: the whole level loop is ONE launch of the persistent slab kernel (in-kernel acquire / push / release)
    out = _run(29543, {"FDW_SAME_DEVICE": "1"}, expect)
    assert counter(out, "persistent slab launches on rank 0:") > 0
    # the same job through the per-level launches replayed as a CUDA graph (what large slabs use)
    out = _run(29545, {"FDW_SAME_DEVICE": "1", "FDW_PSLAB": "0"}, expect)
    assert counter(out, "graph replays on rank 0:") > 0
    if torch.cuda.device_count() >= 2:  # one process per GPU over NVLink, NCCL rendezvous, both halo transports
        _run(29541, {}, ("halo=p2p vs single domain bitwise: OK", "halo=nccl vs single domain bitwise: OK",
                         "mod_main+rtm_main x2 halo=p2p vs one GPU bitwise: OK",
                         "chained stack vs sequential bitwise: OK"))
