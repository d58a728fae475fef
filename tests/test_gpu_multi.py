"""`-m gpu`, needs >= 2 GPUs (skipped otherwise): the sharded path over NCCL against the oracle."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_run_over_nccl_matches_oracle(product_lib, world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(HERE, "dist_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MISMATCH" not in res.stdout


@pytest.mark.parametrize("world", [2, 8])
def test_cpp_host_drives_several_gpus_through_the_c_abi(product_lib, world):
    """tests/cabi/consumer_multi.cpp: planner + device-resident simulate on every GPU +
    modle_b200_reduce_band (ncclReduce) from ONE C++ process, no Python; the reduced band must
    equal the unsharded single-GPU run bit for bit."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import json

    import cabi_build

    exe = cabi_build.build_consumer_multi()
    res = subprocess.run([exe, str(world), "16"], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["ok"] is True and out["world"] == world and out["shards"] >= world
    assert out["band_sum"] + out["missed"] == out["contacts"]
