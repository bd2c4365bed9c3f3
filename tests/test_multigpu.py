"""Multi-GPU correctness on hardware: the batch sharded over N GPUs, forward on
each, logits all-gathered by the library's NVLink collective == the single-GPU logits of the same images, bit for bit.
Needs >= 2 GPUs in one box (`gpurun --gpus N`); spawns torchrun with one rank per GPU. The world_size-2 host logic is
covered on CPU by tests/test_parallel.py (gloo)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("arch,per_rank", [("vit_base", 64), ("resnet50", 32)])
def test_sharded_forward_plus_allgather_equals_single_gpu_bitwise(arch, per_rank):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs in one box")
    env = dict(os.environ, EQXV_TEST_ARCH=arch, EQXV_TEST_PER_RANK=str(per_rank))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert f"MULTIGPU world={n} arch={arch}" in r.stdout and "bitwise_equal=True" in r.stdout, r.stdout[-2000:]
    print(r.stdout.strip().splitlines()[-1])
