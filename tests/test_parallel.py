"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a batch, run a stand-in forward and
all-gather the logits; the result must equal the single-process result bit for bit."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_the_batch_exactly():
    from eqxvision_b200.parallel import shard_bounds

    for n in (0, 1, 7, 8, 255, 256, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _worker(rank, world, port, n_images, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from eqxvision_b200 import parallel
    from oracle import checkpoints as ck
    from oracle import models as om

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sd = ck.vit_state_dict(embed_dim=64, depth=1, heads=1, num_classes=5, seed=3)
    images = ck.synthetic_images(n_images, seed=4)
    fwd = lambda x: om.vit(sd, x, heads=1) if x.shape[0] else torch.zeros(0, 5)  # noqa: E731  CPU stand-in
    full = parallel.data_parallel_forward(fwd, images, gather=True)
    local = parallel.data_parallel_forward(fwd, images, gather=False)
    lo, hi = parallel.shard_bounds(n_images, rank, world)
    assert local.shape[0] == hi - lo
    if rank == 0:
        torch.save({"full": full, "ref": fwd(images)}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [4, 5])
def test_two_rank_shard_and_gather_equals_single_process(tmp_path, n_images):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(2, port, n_images, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["full"].shape == (n_images, 5)
    # the CPU stand-in runs torch GEMMs whose blocking (hence fp32 summation order) depends on the shard size
    assert torch.allclose(r["full"], r["ref"], atol=2e-5, rtol=1e-4)


def test_single_process_is_identity():
    from eqxvision_b200 import parallel

    x = torch.arange(12.).reshape(6, 2)
    assert torch.equal(parallel.shard(x), x)
    assert torch.equal(parallel.data_parallel_forward(lambda t: t * 2, x), x * 2)
