"""Multi-GPU numerics on hardware: spawns one process per GPU (NCCL) and runs tests/_dist_worker.py.

Skipped on a box with fewer than 2 GPUs; run with ``gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`` (the log
of such a run is committed under profiles/).  The checks themselves are listed in tests/_dist_worker.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

EXPECTED = ["reduced_gradients_equal_oracle_data_parallel", "reduced_gradients_identical_on_all_ranks", "reduced_gradients_equal_sum_of_shards",
            "replicas_bit_identical_after_4_steps", "federated_allreduce_equals_oracle_fedavg"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_numerics():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           str(_free_port()), os.path.join(HERE, "_dist_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200)
    print(r.stdout[-6000:])
    assert r.returncode == 0, r.stdout[-3000:]
    for name in EXPECTED:
        assert f"DIST_OK {name}" in r.stdout, name
    assert "DIST_ALL_OK" in r.stdout
