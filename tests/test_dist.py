"""Multi-rank path: cell partition, interface plan, interface-row exchange, sharded CG.
CPU: world_size 2 and 3 over gloo (host logic; local systems from the oracle).  GPU: world_size 2 over NCCL through the
C-ABI (needs two GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_dist.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _spawn(world, backend, port, tmp_path):
    out = str(tmp_path / "result")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), str(r), str(world), backend, str(port), out],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(world)]
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(o.decode()[-3000:])
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{logs[r]}"
        assert open(out + f".{r}").read().startswith("ok")
    return [open(out + f".{r}").read() for r in range(world)]


@pytest.mark.parametrize("world", [2, 3])
def test_partition_exchange_gloo(world, tmp_path):
    _spawn(world, "gloo", 29611 + world, tmp_path)


@pytest.mark.gpu
def test_partition_exchange_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    print(_spawn(2, "nccl", 29621, tmp_path))
